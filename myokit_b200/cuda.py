"""
Device information and selection for the CUDA back-end: the counterpart of
``myokit.OpenCL`` (``myokit/_sim/opencl.py:19``, backed by ``mcl.h:265,823``).

OpenCL needs a platform / device pair and a ``preferred-opencl-device.ini``;
here a device is a CUDA ordinal. The preferred ordinal can be stored in
``~/.config/myokit/preferred-cuda-device.ini`` (``save_selection``) and is what
``SimulationCUDA(device=None)`` uses.
"""
import configparser
import os

from . import capi

_INI = 'preferred-cuda-device.ini'


class NoCUDAError(Exception):
    """Raised when information is requested but no CUDA device is usable."""


class CUDADeviceInfo:
    """Properties of one device (cf. ``OpenCLDeviceInfo``, ``opencl.py:366``)."""

    def __init__(self, ordinal, info):
        self.ordinal = ordinal
        self.name = info['name']
        self.compute_capability = info['compute_capability']
        self.sm_count = info['sm_count']
        self.clock_khz = info['clock_khz']
        self.total_mem = info['total_mem']
        self.l2_bytes = info['l2_bytes']
        self.smem_per_block_optin = info['smem_per_block_optin']

    def is_blackwell(self):
        """True for compute capability 10.x (the kernels target sm_100a)."""
        return self.compute_capability[0] == 10

    def format(self, pre=''):
        lines = [
            pre + 'Device ' + str(self.ordinal) + ': ' + self.name,
            pre + ' Compute capability: %d.%d' % self.compute_capability
            + ('' if self.is_blackwell() else '  (unsupported: sm_100a only)'),
            pre + ' Multiprocessors   : ' + str(self.sm_count),
            pre + ' Clock speed       : ' + clockspeed(self.clock_khz * 1e3),
            pre + ' Global memory     : ' + bytesize(self.total_mem),
            pre + ' L2 cache          : ' + bytesize(self.l2_bytes),
            pre + ' Shared mem / block: ' + bytesize(self.smem_per_block_optin),
        ]
        return '\n'.join(lines)


class CUDA:
    """Static information methods, mirroring ``myokit.OpenCL``."""

    @staticmethod
    def supported():
        """True if the native library loads and sees at least one device."""
        try:
            return capi.device_count() > 0
        except Exception:
            return False

    @staticmethod
    def available():
        """Returns a list of :class:`CUDADeviceInfo`, one per device."""
        return [CUDADeviceInfo(i, capi.device_info(i))
                for i in range(capi.device_count())]

    @staticmethod
    def info(formatted=False):
        """All devices: objects, or a formatted string."""
        devices = CUDA.available()
        if not formatted:
            return devices
        if not devices:
            return 'No CUDA devices found.'
        sel = CUDA.load_selection()
        out = []
        for d in devices:
            text = d.format()
            if d.ordinal == sel:
                text += '\n (selected)'
            out.append(text)
        return ('\n' + '-' * 60 + '\n').join(out)

    @staticmethod
    def current_info(formatted=False):
        """Information about the device simulations will use."""
        n = capi.device_count()
        if n == 0:
            raise NoCUDAError('No CUDA devices found.')
        sel = CUDA.load_selection()
        if sel >= n:
            sel = 0
        d = CUDADeviceInfo(sel, capi.device_info(sel))
        return d.format() if formatted else d

    @staticmethod
    def _path():
        try:
            import myokit
            base = myokit.DIR_USER
        except Exception:   # pragma: no cover
            base = os.path.join(os.path.expanduser('~'), '.config', 'myokit')
        return os.path.join(base, _INI)

    @staticmethod
    def load_selection():
        """Preferred device ordinal (0 if nothing was saved)."""
        env = os.environ.get('MYOKIT_CUDA_DEVICE')
        if env is not None:
            try:
                return int(env)
            except ValueError:
                pass
        path = CUDA._path()
        if os.path.isfile(path):
            c = configparser.ConfigParser()
            try:
                c.read(path)
                return c.getint('selection', 'device')
            except Exception:
                return 0
        return 0

    @staticmethod
    def save_selection(device=None):
        """Stores the preferred device ordinal (``None`` removes it)."""
        path = CUDA._path()
        if device is None:
            if os.path.isfile(path):
                os.remove(path)
            return
        c = configparser.ConfigParser()
        c.add_section('selection')
        c.set('selection', 'device', str(int(device)))
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, 'w') as f:
            c.write(f)


def bytesize(size):
    """Human-readable byte count."""
    size = float(size)
    for unit in ('B', 'KB', 'MB', 'GB', 'TB'):
        if size < 1024 or unit == 'TB':
            return ('%.4g %s' % (size, unit))
        size /= 1024


def clockspeed(hz):
    """Human-readable clock speed."""
    hz = float(hz)
    for unit in ('Hz', 'kHz', 'MHz', 'GHz'):
        if hz < 1000 or unit == 'GHz':
            return ('%.4g %s' % (hz, unit))
        hz /= 1000
