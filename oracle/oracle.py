"""
CPU oracle for the SimulationOpenCL hot path (TEST INFRASTRUCTURE ONLY).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs may import this. Nothing under ``myokit_b200/`` does, and the product
never falls back to it.

``OracleSimulation`` has the constructor and setters of the reference's
``SimulationOpenCL`` (``myokit/_sim/openclsim.py:149``) and a ``run`` that
returns ``(log, state)`` as numpy data. Two kernels can sit under the same
driver (``oracle/driver.c``):

``kernel='port'``
    ``oracle/cgen.py``: this repo's restatement of the reference's generator.
``kernel='ref'``
    The reference's own ``myokit/_sim/openclsim.cl`` rendered by the
    reference's own template engine with the argument dict of
    ``openclsim.py:1060-1073`` and compiled as C behind ``oracle/cl_shim.h``.
    Outputs go to ``oracle/_ref/`` (git-ignored). Needs the reference
    package (``baseline/_ref`` or ``/root/reference``) at BUILD time only.

Parity status: pinned — see ``tests/test_oracle.py``: the port is checked
against the reference's ``Simulation1d`` (run in the build container; golden
vectors in ``tests/golden``), against ``kernel='ref'``, and against the
cross-implementation tolerances of
``myokit/tests/test_simulation_opencl_vs_sim1d.py:118-136``.
"""
import ctypes
import time
import hashlib
import io
import os
import subprocess

import numpy as np

from ._locate import import_myokit
from . import cgen

myokit = import_myokit()

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, '_build')
_REF = os.path.join(_HERE, '_ref')

LOG_TIME, LOG_PACE, LOG_IDIFF, LOG_STATE, LOG_INTER = range(5)


def _host_isa_tag():
    try:
        with open('/proc/cpuinfo') as f:
            for line in f:
                if line.startswith('flags'):
                    return hashlib.sha1(line.encode()).hexdigest()[:12]
    except OSError:
        pass
    return 'unknown'


def _compile(header_text, tag, ref_kernel, openmp, contract, opt):
    """Compiles driver.c against a model header; returns the .so path."""
    outdir = _REF if ref_kernel else _BUILD
    os.makedirs(outdir, exist_ok=True)
    with open(os.path.join(_HERE, 'driver.c'), 'rb') as f:
        drv = f.read()
    with open(os.path.join(_HERE, 'cl_shim.h'), 'rb') as f:
        shim = f.read()
    flags = [opt, '-fPIC', '-shared', '-std=gnu11',
             '-ffp-contract=' + ('fast' if contract else 'off'), '-w']
    if opt == '-O3':
        # Timed CPU-baseline builds; the cache key must not let a binary
        # built for another host's ISA be reused.
        flags.append('-march=native')
        flags.append('-DORACLE_HOST_ISA_%s=1' % _host_isa_tag())
    if openmp:
        flags.append('-fopenmp')
    h = hashlib.sha1()
    for part in (header_text.encode(), drv, shim, ' '.join(flags).encode()):
        h.update(part)
    name = '%s_%s' % (tag, h.hexdigest()[:16])
    so = os.path.join(outdir, name + '.so')
    if os.path.isfile(so):
        return so
    hdr = os.path.join(outdir, name + '.h')
    with open(hdr, 'w') as f:
        f.write(header_text)
    cmd = ['gcc'] + flags + [
        '-I', _HERE, '-DORACLE_MODEL_HEADER="%s"' % hdr,
    ]
    if ref_kernel:
        cmd.append('-DORACLE_REF_KERNEL=1')
    tmp = so + '.tmp%d' % os.getpid()
    cmd += [os.path.join(_HERE, 'driver.c'), '-o', tmp, '-lm']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('Oracle compilation failed:\n' + r.stderr)
    os.replace(tmp, so)
    return so


def render_reference_kernel(model, precision, bound_variables, inter_log,
                            diffusion, fields, paced_cells, rl_states,
                            connections, heterogeneous, fiber_tissue=False):
    """
    Renders the REFERENCE's ``openclsim.cl`` with the reference's own template
    engine, argument dict as in ``myokit/_sim/openclsim.py:1060-1073``.
    """
    import myokit.pype
    args = {
        'model': model,
        'precision': precision,
        'native_math': False,
        'bound_variables': bound_variables,
        'inter_log': inter_log,
        'diffusion': diffusion,
        'fields': fields,
        'paced_cells': paced_cells,
        'rl_states': rl_states,
        'connections': connections,
        'heterogeneous': heterogeneous,
        'fiber_tissue': bool(fiber_tissue),
    }
    e = myokit.pype.TemplateEngine()
    s = io.StringIO()
    e.set_output_stream(s)
    e.process(os.path.join(myokit.DIR_CFUNC, 'openclsim.cl'), args)
    text = s.getvalue()
    pre = ['/* Rendered from the reference\'s myokit/_sim/openclsim.cl */']
    post = ['#define N_STATE n_state', '#define N_INTER n_inter',
            '#define N_FIELD n_field']
    post.append('#define I_VM i_vm' if diffusion else '#define I_VM 0')
    if fiber_tissue:
        post.append('#define ORACLE_REF_FT 1')
    if not (diffusion and connections):
        post.append('#define diff_arb_reset(a, b) ((void)0)')
        post.append('#define diff_arb_step(a, b, c, d, e, f) ((void)0)')
    if not heterogeneous:
        post.append('#define diff_hetero(a, b, c, d, e, f) ((void)0)')
    if not (diffusion and not connections and not heterogeneous):
        post.append('#define diff_step(a, b, c, d, e, f) ((void)0)')
    return '\n'.join(pre) + '\n' + text + '\n' + '\n'.join(post) + '\n'


class OracleSimulation:
    """
    CPU oracle with the ``SimulationOpenCL`` surface (subset needed by tests).
    """

    def __init__(self, model, protocol=None, ncells=256, diffusion=True,
                 precision=None, rl=False, kernel='port', openmp=False,
                 contract=False, opt='-O2'):
        if precision is None:
            precision = myokit.SINGLE_PRECISION
        try:
            self._nx, self._ny = int(ncells), 1
            self._dims = (self._nx,)
        except TypeError:
            self._nx, self._ny = int(ncells[0]), int(ncells[1])
            self._dims = (self._nx, self._ny)
        self._n = self._nx * self._ny
        self._diffusion = bool(diffusion)
        self._precision = precision
        self._kernel = kernel
        self._openmp = bool(openmp)
        self._contract = bool(contract)
        self._opt = opt
        self._protocol = None if protocol is None else protocol.clone()
        (self._model, self._vm, self._rl_states,
         self._bound) = cgen.prepare_model(model, self._diffusion, rl)
        self._nstate = self._model.count_states()
        self._gx, self._gy = 10.0, 5.0
        self._gx_field = self._gy_field = None
        self._connections = None
        self._paced = (5, 5, 0, 0) if len(self._dims) == 2 else (5, 1, 0, 0)
        self._fields = {}
        self._dt = 0.005
        self._time = 0.0
        self._state = np.tile(np.array(
            self._model.initial_values(True), dtype=np.float64), self._n)
        self.last_steps = 0
        self.last_run_seconds = 0.0     # wall clock of the native run call
        self.last_loop_seconds = 0.0    # ... of its time-step loop alone
        self.last_halted = False

    # -- setters (subset of openclsim.py:1286-1715) --------------------------
    def set_conductance(self, gx=10, gy=5):
        self._gx, self._gy = float(gx), float(gy)
        self._gx_field = self._gy_field = None
        self._connections = None

    def set_conductance_field(self, gx, gy=None):
        self._gx_field = np.ascontiguousarray(gx, dtype=np.float64).ravel()
        if gy is None:
            self._gy_field = np.zeros(0)
        else:
            self._gy_field = np.ascontiguousarray(
                gy, dtype=np.float64).ravel()
        self._connections = None

    def set_connections(self, connections):
        conns = []
        for i, j, c in connections:
            i, j = int(i), int(j)
            i, j = (i, j) if i < j else (j, i)
            conns.append((i, j, float(c)))
        self._connections = conns
        self._gx_field = self._gy_field = None

    def set_paced_cells(self, nx=5, ny=5, x=0, y=0):
        # openclsim.py:1569-1593
        nx, x = int(nx), int(x)
        if nx < 0:
            nx = -nx
            x -= nx
        if x < 0:
            x += self._nx
        if len(self._dims) == 1:
            ny, y = 1, 0
        else:
            ny, y = int(ny), int(y)
            if ny < 0:
                ny = -ny
                y -= ny
            if y < 0:
                y += self._ny
        self._paced = (nx, ny, x, y)

    def set_paced_cell_list(self, cells):
        if len(self._dims) == 1:
            self._paced = [int(c) for c in cells]
        else:
            self._paced = [int(i) + int(j) * self._nx for i, j in cells]

    def set_field(self, var, values):
        if isinstance(var, myokit.Variable):
            var = var.qname()
        var = self._model.get(var)
        self._fields[var] = np.ascontiguousarray(
            values, dtype=np.float64).reshape(self._n)

    def set_constant(self, var, value):
        if isinstance(var, myokit.Variable):
            var = var.qname()
        self._model.set_value(self._model.get(var).qname(), float(value))

    def set_step_size(self, dt=0.005):
        self._dt = float(dt)

    def set_state(self, state):
        state = np.asarray(state, dtype=np.float64).ravel()
        if state.size == self._nstate:
            state = np.tile(state, self._n)
        self._state = state.copy()

    def set_time(self, t=0):
        self._time = float(t)

    def set_protocol(self, protocol=None):
        self._protocol = None if protocol is None else protocol.clone()

    def state(self):
        return self._state.copy()

    def time(self):
        return self._time

    # -- helpers --------------------------------------------------------------
    def _paced_mask(self):
        mask = np.zeros(self._n, dtype=np.uint8)
        if isinstance(self._paced, tuple):
            nx, ny, x, y = self._paced
            xs = np.arange(self._nx)
            ys = np.arange(self._ny)
            mx = (xs >= x) & (xs < x + nx)
            my = (ys >= y) & (ys < y + ny)
            mask = (my[:, None] & mx[None, :]).astype(np.uint8).ravel()
        else:
            for cid in self._paced:
                mask[cid] = 1
        return np.ascontiguousarray(mask)

    def _events(self):
        if self._protocol is None:
            return np.zeros(0)
        ev = []
        for e in self._protocol.events():
            ev.extend([e.level(), e.start(), e.duration(), e.period(),
                       e.multiplier()])
        return np.array(ev, dtype=np.float64)

    def _library(self, inter_log):
        fields = list(self._fields.keys())
        tag = 'oracle_%s_%d' % (self._kernel, self._precision)
        if self._kernel == 'ref':
            text = render_reference_kernel(
                self._model, self._precision, self._bound, inter_log,
                self._diffusion, fields, self._paced, self._rl_states,
                self._connections is not None, self._gx_field is not None)
        else:
            text = cgen.generate(
                self._model, self._precision, self._bound, inter_log,
                self._diffusion, fields, self._rl_states)
        so = _compile(text, tag, self._kernel == 'ref', self._openmp,
                      self._contract, self._opt)
        lib = ctypes.CDLL(so)
        lib.oracle_run.restype = ctypes.c_int
        return lib

    def run(self, duration, log=None, log_interval=1.0, nthreads=0):
        """
        Runs; returns ``(log, state)`` with ``log`` a dict
        ``key -> numpy array`` keyed like the reference's DataLog.
        """
        tmin = self._time
        tmax = tmin + duration

        # Log preparation: openclsim.py:1025-1053
        g = []
        for label in ('time', 'pace'):
            var = self._model.binding(label)
            if var is not None:
                g.append(var.qname())
        dlog = myokit.prepare_log(
            log, self._model, dims=self._dims, global_vars=g,
            if_empty=myokit.LOG_STATE + myokit.LOG_BOUND,
            allowed_classes=myokit.LOG_STATE + myokit.LOG_INTER
            + myokit.LOG_BOUND, precision=self._precision)
        inter_log = []
        seen = set()
        for key in dlog.keys():
            name = myokit.split_key(key)[1]
            if name in seen:
                continue
            seen.add(name)
            var = self._model.get(name)
            if var.is_intermediary() and not var.is_bound():
                inter_log.append(var)
        inter_index = dict((v.qname(), k) for k, v in enumerate(inter_log))
        n_inter = len(inter_log)

        log_interval = 1e-9 if log_interval is None else float(log_interval)
        if log_interval <= 0:
            log_interval = 1e-9

        keys = list(dlog.keys())
        kinds = np.zeros(len(keys), dtype=np.int32)
        index = np.zeros(len(keys), dtype=np.uint64)
        vtime = self._model.binding('time')
        vpace = self._model.binding('pace')
        vdiff = self._model.binding('diffusion_current') \
            if self._diffusion else None
        for i, key in enumerate(keys):
            cell, name = myokit.split_key(key)
            var = self._model.get(name)
            if cell == '':
                if var is vtime:
                    kinds[i] = LOG_TIME
                elif var is vpace:
                    kinds[i] = LOG_PACE
                else:
                    raise ValueError('Unknown global ' + key)
                continue
            parts = [int(x) for x in cell.split('.') if x != '']
            cid = parts[0] + (parts[1] * self._nx if len(parts) > 1 else 0)
            if var is vdiff:
                kinds[i], index[i] = LOG_IDIFF, cid
            elif var.is_state():
                kinds[i] = LOG_STATE
                index[i] = cid * self._nstate + var.index()
            elif name in inter_index:
                kinds[i] = LOG_INTER
                index[i] = cid * n_inter + inter_index[name]
            else:
                raise ValueError('Cannot log ' + key)

        state = np.ascontiguousarray(self._state, dtype=np.float64).copy()
        if duration <= 0:
            return dict((k, np.zeros(0)) for k in keys), state

        lib = self._library(inter_log)
        fields = list(self._fields.values())
        if fields:
            # cell-major: field_data[cid * n_field + k] (openclsim.py:1082-1089)
            field_data = np.ascontiguousarray(
                np.vstack(fields).T, dtype=np.float64).ravel()
        else:
            field_data = np.zeros(1)

        mode = 0
        gxf = gyf = np.zeros(1)
        c1 = c2 = np.zeros(1, dtype=np.uint64)
        cg = np.zeros(1)
        n_conn = 0
        if self._diffusion:
            mode = 1
            if self._connections is not None:
                mode = 3
                n_conn = len(self._connections)
                c1 = np.array([c[0] for c in self._connections], np.uint64)
                c2 = np.array([c[1] for c in self._connections], np.uint64)
                cg = np.array([c[2] for c in self._connections], np.float64)
            elif self._gx_field is not None:
                mode = 2
                gxf = np.concatenate([self._gx_field, np.zeros(1)])
                gyf = np.concatenate([self._gy_field, np.zeros(1)])
        mask = self._paced_mask() if self._diffusion else np.zeros(
            1, dtype=np.uint8)
        events = self._events()
        n_events = len(events) // 5
        if n_events == 0:
            events = np.zeros(5)

        # Upper bound on log rows
        max_rows = int(duration / log_interval) + 16
        max_rows = min(max_rows, int(duration / self._dt) * 2 + 16)
        out = np.zeros((max_rows, max(len(keys), 1)), dtype=np.float64)
        n_rows = ctypes.c_uint64(0)
        n_steps = ctypes.c_uint64(0)
        halted = ctypes.c_int(0)
        tfinal = ctypes.c_double(0)

        def ptr(a, t):
            return a.ctypes.data_as(ctypes.POINTER(t))

        t_call = time.perf_counter()
        rc = lib.oracle_run(
            ctypes.c_size_t(self._nx), ctypes.c_size_t(self._ny),
            ctypes.c_int(mode),
            ctypes.c_double(self._gx), ctypes.c_double(self._gy),
            ptr(gxf, ctypes.c_double), ptr(gyf, ctypes.c_double),
            ctypes.c_size_t(n_conn), ptr(c1, ctypes.c_uint64),
            ptr(c2, ctypes.c_uint64), ptr(cg, ctypes.c_double),
            ctypes.c_double(tmin), ctypes.c_double(tmax),
            ctypes.c_double(self._dt), ctypes.c_double(log_interval),
            ptr(state, ctypes.c_double), ptr(field_data, ctypes.c_double),
            ptr(mask, ctypes.c_ubyte),
            ctypes.c_int(n_events), ptr(events, ctypes.c_double),
            ctypes.c_size_t(len(keys)), ptr(kinds, ctypes.c_int),
            ptr(index, ctypes.c_uint64),
            ptr(out, ctypes.c_double), ctypes.c_size_t(max_rows),
            ctypes.byref(n_rows), ctypes.byref(n_steps),
            ctypes.byref(halted), ctypes.byref(tfinal),
            ctypes.c_int(nthreads))
        self.last_run_seconds = time.perf_counter() - t_call
        try:
            lib.oracle_loop_seconds.restype = ctypes.c_double
            self.last_loop_seconds = float(lib.oracle_loop_seconds())
        except AttributeError:      # a library built before the timer existed
            self.last_loop_seconds = self.last_run_seconds
        if rc < 0:
            raise RuntimeError('Oracle pacing error %d' % rc)
        if rc > 0:
            raise RuntimeError('Oracle log buffer overflow')
        self.last_steps = int(n_steps.value)
        self.last_halted = bool(halted.value)
        rows = int(n_rows.value)
        result = dict((k, out[:rows, i].copy()) for i, k in enumerate(keys))
        self._state = state
        self._time = tmax
        return result, state.copy()


def pacing_probe(events, times, t0=0.0):
    """
    Runs the oracle's pacing restatement alone; returns ``(levels, tnexts)``
    after advancing to each of ``times``. ``events`` is a list of
    ``(level, start, duration, period, multiplier)``.
    """
    text = ('#define N_STATE 1\n#define N_INTER 0\n#define N_FIELD 0\n'
            '#define I_VM 0\ntypedef double Real;\n'
            'static void cell_step(const size_t cid, const Real time, '
            'const Real dt, const Real pace, Real* state, '
            'const Real* idiff_in, Real* inter_log, const Real* field_data)'
            '{ (void)cid; }\n')
    so = _compile(text, 'oracle_pacing', False, False, False, '-O2')
    lib = ctypes.CDLL(so)
    ev = np.array(events, dtype=np.float64).ravel()
    n_events = len(ev) // 5
    if n_events == 0:
        ev = np.zeros(5)
    times = np.ascontiguousarray(times, dtype=np.float64)
    levels = np.zeros(len(times))
    tnexts = np.zeros(len(times))
    dp = ctypes.POINTER(ctypes.c_double)
    rc = lib.oracle_pacing_probe(
        ctypes.c_double(t0), ctypes.c_int(n_events), ev.ctypes.data_as(dp),
        ctypes.c_int(len(times)), times.ctypes.data_as(dp),
        levels.ctypes.data_as(dp), tnexts.ctypes.data_as(dp))
    if rc:
        raise RuntimeError('Oracle pacing error %d' % rc)
    return levels, tnexts
