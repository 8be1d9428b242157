"""
ctypes binding of ``libmyokit_b200.so`` (``include/myokit_b200.h``).

The library is the product: if it cannot be loaded, importing this module
raises — there is no Python or CPU fallback.
"""
import ctypes
import os

import numpy as np

from . import build as _build

c_u64 = ctypes.c_uint64
c_i64 = ctypes.c_int64
c_vp = ctypes.c_void_p

MKB_ABI_VERSION = 5
MKB_OK = 0
MKB_ERR_INVALID = -1
MKB_ERR_CUDA = -2
MKB_ERR_JIT = -3
MKB_ERR_PACING = -4
MKB_ERR_SIMULTANEOUS = -5
MKB_ERR_STATE = -6

LOG_TIME, LOG_PACE, LOG_IDIFF, LOG_STATE, LOG_INTER = range(5)
LOG_STATE_FIELD, LOG_INTER_FIELD, LOG_IDIFF_FIELD = 5, 6, 7

# Every symbol include/myokit_b200.h declares
SYMBOLS = [
    'mkb_abi_version', 'mkb_host_alloc', 'mkb_host_free', 'mkb_last_error', 'mkb_free', 'mkb_device_count',
    'mkb_device_info', 'mkb_device_abi_header', 'mkb_jit_compile',
    'mkb_sim_init', 'mkb_sim_step', 'mkb_sim_log_view', 'mkb_sim_get_state',
    'mkb_sim_counters', 'mkb_sim_device_ms', 'mkb_sim_set_steps_per_call',
    'mkb_sim_reset_counters', 'mkb_sim_clean', 'mkb_sim_halo_info',
    'mkb_sim_halo_export', 'mkb_sim_halo_connect', 'mkb_sim_halo_seed',
    'mkb_sim_rearm', 'mkb_sim_set_state', 'mkb_sim_halo_live', 'mkb_measure_peaks', 'mkb_sim_ghost_connect',
    'mkb_pacing_probe', 'mkb_schedule_probe',
    'mkb_sim_junction_connect', 'mkb_sim_step_pair',
]


class DeviceInfo(ctypes.Structure):
    _fields_ = [
        ('name', ctypes.c_char * 256),
        ('cc_major', ctypes.c_int), ('cc_minor', ctypes.c_int),
        ('sm_count', ctypes.c_int), ('clock_khz', ctypes.c_int),
        ('total_mem', ctypes.c_size_t), ('l2_bytes', ctypes.c_size_t),
        ('smem_per_block_optin', ctypes.c_size_t),
    ]


class SimConfig(ctypes.Structure):
    _fields_ = [
        ('abi_version', ctypes.c_int), ('device', ctypes.c_int),
        ('precision', ctypes.c_int), ('host_precision', ctypes.c_int),
        ('cubin', c_vp), ('cubin_size', ctypes.c_size_t),
        ('kernel_name', ctypes.c_char_p),
        ('block_x', ctypes.c_int), ('block_y', ctypes.c_int),
        ('cells_per_thread', ctypes.c_int), ('rows_per_thread', ctypes.c_int),
        ('n_state', ctypes.c_int), ('i_vm', ctypes.c_int),
        ('n_inter', ctypes.c_int), ('n_field', ctypes.c_int),
        ('nx', c_u64), ('ny', c_u64),
        ('diffusion_mode', ctypes.c_int),
        ('gx', ctypes.c_double), ('gy', ctypes.c_double),
        ('gx_field', c_vp), ('gy_field', c_vp),
        ('n_ghost', c_u64),
        ('n_connections', c_u64), ('conn_i', c_vp), ('conn_j', c_vp),
        ('conn_g', c_vp),
        ('pace_rect', ctypes.c_int),
        ('pace_nx', c_i64), ('pace_ny', c_i64),
        ('pace_x', c_i64), ('pace_y', c_i64),
        ('n_paced', c_u64), ('paced_cells', c_vp),
        ('n_events', ctypes.c_int), ('events', c_vp),
        ('tmin', ctypes.c_double), ('tmax', ctypes.c_double),
        ('dt', ctypes.c_double), ('log_interval', ctypes.c_double),
        ('state_in', c_vp), ('field_data', c_vp),
        ('n_log', c_u64), ('log_kind', c_vp), ('log_index', c_vp),
        ('iy_offset', c_u64), ('ny_global', c_u64),
        ('steps_per_call', c_u64), ('use_graphs', ctypes.c_int),
        ('state_uniform', ctypes.c_int),
        ('second_kernel_name', ctypes.c_char_p),
        ('kernel_flags', ctypes.c_int),
        ('stream_box_w', ctypes.c_int), ('stream_box_h', ctypes.c_int),
        ('kernel_smem_bytes', ctypes.c_int),
        ('kernel_stride', c_u64),
    ]


KERNEL_PERSISTENT = 1
KERNEL_STREAM = 2
KERNEL_OVERLAP = 4
KERNEL_STAGE = 8
KERNEL_TILE_LOOP = 16


class RunConfig(ctypes.Structure):
    _fields_ = [
        ('tmin', ctypes.c_double), ('tmax', ctypes.c_double),
        ('dt', ctypes.c_double), ('log_interval', ctypes.c_double),
        ('n_events', ctypes.c_int), ('events', c_vp),
        ('n_log', c_u64), ('log_kind', c_vp), ('log_index', c_vp),
        ('steps_per_call', c_u64),
    ]


class GhostPeer(ctypes.Structure):
    _fields_ = [
        ('handle', c_vp), ('peer_n_ghost', c_u64),
        ('peer_n_flags', ctypes.c_uint32), ('flag_index', ctypes.c_uint32),
        ('n_export', c_u64), ('src_cell', c_vp), ('dst_slot', c_vp),
    ]


class _PinnedBlock:
    """Page-locked host memory from mkb_host_alloc, freed with the last view."""
    def __init__(self, nbytes):
        lib = library()
        ptr = c_vp()
        check(lib.mkb_host_alloc(ctypes.c_size_t(nbytes), ctypes.byref(ptr)))
        self.ptr, self.nbytes, self._free = ptr.value, nbytes, lib.mkb_host_free

    def __del__(self):
        try:
            self._free(c_vp(self.ptr))
        except Exception:   # pragma: no cover
            pass


def host_array(count, dtype=np.float64, pinned_from=1 << 20):
    """
    An uninitialised 1-d host array; page-locked (copied to and from the
    device by asynchronous DMA at PCIe speed) when it is at least
    ``pinned_from`` bytes and a CUDA device is there, ordinary memory otherwise.
    """
    dtype = np.dtype(dtype)
    nbytes = int(count) * dtype.itemsize
    if nbytes >= pinned_from:
        try:
            blk = _PinnedBlock(nbytes)
        except BackendError:
            blk = None
        if blk is not None:
            buf = (ctypes.c_char * nbytes).from_address(blk.ptr)
            buf._mkb_block = blk        # the array keeps buf, buf keeps blk
            return np.frombuffer(buf, dtype=dtype, count=int(count))
    return np.empty(int(count), dtype=dtype)


class BackendError(Exception):
    """Raised when the native library reports an error."""
    def __init__(self, code, message):
        super().__init__(message)
        self.code = code


_lib = None


def library():
    """Loads (building if needed) and returns the native library."""
    global _lib
    if _lib is not None:
        return _lib
    # (stamp check: a library built from older sources is rebuilt, never
    # loaded next to a newer generator)
    path = _build.build_library(force=False)
    lib = ctypes.CDLL(path)
    lib.mkb_abi_version.restype = ctypes.c_int
    lib.mkb_last_error.restype = ctypes.c_char_p
    lib.mkb_free.argtypes = [c_vp]
    lib.mkb_free.restype = None
    lib.mkb_host_alloc.argtypes = [ctypes.c_size_t, ctypes.POINTER(c_vp)]
    lib.mkb_host_free.argtypes = [c_vp]
    lib.mkb_host_free.restype = None
    lib.mkb_device_count.restype = ctypes.c_int
    lib.mkb_device_info.argtypes = [ctypes.c_int, ctypes.POINTER(DeviceInfo)]
    lib.mkb_device_abi_header.restype = ctypes.c_char_p
    lib.mkb_jit_compile.argtypes = [
        ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(c_vp),
        ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(c_vp)]
    lib.mkb_sim_init.argtypes = [ctypes.POINTER(SimConfig), ctypes.POINTER(c_vp)]
    lib.mkb_sim_step.argtypes = [
        c_vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)]
    lib.mkb_sim_log_view.argtypes = [
        c_vp, ctypes.POINTER(c_vp), ctypes.POINTER(c_u64),
        ctypes.POINTER(c_u64), ctypes.POINTER(c_u64)]
    lib.mkb_sim_get_state.argtypes = [c_vp, c_vp]
    lib.mkb_sim_counters.argtypes = [
        c_vp, ctypes.POINTER(c_u64), ctypes.POINTER(c_u64)]
    lib.mkb_sim_device_ms.argtypes = [c_vp, ctypes.POINTER(ctypes.c_double)]
    lib.mkb_sim_set_steps_per_call.argtypes = [c_vp, c_u64]
    lib.mkb_sim_reset_counters.argtypes = [c_vp]
    lib.mkb_sim_clean.argtypes = [c_vp]
    lib.mkb_sim_clean.restype = None
    lib.mkb_sim_halo_info.argtypes = [
        c_vp, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
        ctypes.POINTER(c_u64)]
    lib.mkb_sim_halo_export.argtypes = [c_vp, c_vp, ctypes.POINTER(c_vp)]
    lib.mkb_sim_halo_connect.argtypes = [c_vp, c_vp, c_vp, ctypes.c_int]
    lib.mkb_sim_halo_seed.argtypes = [c_vp]
    lib.mkb_sim_set_state.argtypes = [c_vp, c_vp, ctypes.c_int]
    lib.mkb_sim_halo_live.argtypes = [c_vp, ctypes.POINTER(ctypes.c_int)]
    lib.mkb_sim_ghost_connect.argtypes = [
        c_vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(GhostPeer),
        ctypes.c_int, ctypes.c_uint32, c_vp]
    lib.mkb_sim_rearm.argtypes = [c_vp, ctypes.POINTER(RunConfig)]
    lib.mkb_sim_junction_connect.argtypes = [c_vp, c_vp, ctypes.c_double, c_u64]
    lib.mkb_sim_step_pair.argtypes = [
        c_vp, c_vp, c_u64, ctypes.POINTER(ctypes.c_double),
        ctypes.POINTER(ctypes.c_int)]
    lib.mkb_schedule_probe.argtypes = [
        ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double,
        ctypes.c_int, c_vp, c_u64, c_vp, c_vp, c_vp, c_vp,
        ctypes.POINTER(c_u64)]
    lib.mkb_pacing_probe.argtypes = [
        ctypes.c_double, ctypes.c_int, c_vp, ctypes.c_int, c_vp, c_vp, c_vp]
    if lib.mkb_abi_version() != MKB_ABI_VERSION:
        raise ImportError('libmyokit_b200.so ABI version mismatch; rebuild.')
    _lib = lib
    return lib


def check(rc):
    """Raises BackendError for negative return codes."""
    if rc < 0:
        msg = library().mkb_last_error()
        raise BackendError(rc, msg.decode('utf-8', 'replace') if msg else
                           'error %d' % rc)
    return rc


def jit_compile(source, options=()):
    """CUDA source -> sm_100a cubin bytes (NVRTC inside the library)."""
    lib = library()
    cubin = c_vp()
    size = ctypes.c_size_t(0)
    log = c_vp()
    opts = None
    if options:
        opts = b'\0'.join(o.encode('utf-8') for o in options) + b'\0\0'
    rc = lib.mkb_jit_compile(source.encode('utf-8'), opts, ctypes.byref(cubin),
                             ctypes.byref(size), ctypes.byref(log))
    logtext = ''
    if log.value:
        logtext = ctypes.string_at(log.value).decode('utf-8', 'replace')
        lib.mkb_free(log)
    if rc < 0:
        msg = lib.mkb_last_error().decode('utf-8', 'replace')
        raise BackendError(rc, msg)
    data = ctypes.string_at(cubin.value, size.value)
    lib.mkb_free(cubin)
    return data, logtext


def measure_peaks(device=0):
    """Pipe and copy peaks measured on the device (see the C header)."""
    out = (ctypes.c_double * 6)()
    lib = library()
    lib.mkb_measure_peaks.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
    check(lib.mkb_measure_peaks(device, out))
    return {
        'fp64_fma_ginstr_s': out[0], 'fp32_fma_ginstr_s': out[1],
        'mufu_ex2_gop_s': out[2], 'copy_gb_s': out[3],
        'sm_clock_mhz': out[4], 'sm_count': int(out[5]),
    }


def device_count():
    n = library().mkb_device_count()
    return max(n, 0)


def device_info(device=0):
    info = DeviceInfo()
    check(library().mkb_device_info(device, ctypes.byref(info)))
    return {
        'name': info.name.decode('utf-8', 'replace'),
        'compute_capability': (info.cc_major, info.cc_minor),
        'sm_count': info.sm_count,
        'clock_khz': info.clock_khz,
        'total_mem': info.total_mem,
        'l2_bytes': info.l2_bytes,
        'smem_per_block_optin': info.smem_per_block_optin,
    }
