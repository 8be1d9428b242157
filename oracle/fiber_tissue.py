"""
CPU oracle of the fibre-tissue simulation — TEST INFRASTRUCTURE ONLY.

Restates ``myokit.FiberTissueSimulation`` (``myokit/_sim/fiber_tissue.py:17``):
a 2-d fibre and a 2-d tissue, two models, one pacing protocol applied to the
first ``nx_paced`` columns of the fibre; the last fibre column is coupled to
tissue cells ``(0, cty + y)`` with conductance ``g_fiber_tissue``
(``openclsim.cl:601-628``). Forward Euler, homogeneous conductances.

The time loop (``myokit/_sim/fiber_tissue.c:1001-1155``: dt selection with
``dt_min`` = 0, diffusion f / t / junction, state snapshot when logging, cell
steps f / t, log row for time t, pacing advance) is restated here in Python;
every numerical piece is a call into one of two oracle libraries built by
``oracle/oracle.py`` (``kernel='port'``: this repository's restatement;
``kernel='ref'``: the reference's own rendered ``openclsim.cl`` compiled as C,
including its ``diff_step_fiber_tissue``).

Pinned: ``tests/test_oracle_fiber_tissue.py`` — port == reference-rendered
kernels bit for bit; with ``g_fiber_tissue = 0`` both parts equal the
(separately pinned) single-grid oracle bit for bit.
"""
import ctypes

import numpy as np

from ._locate import import_myokit
myokit = import_myokit()

from . import cgen      # noqa: E402
from .oracle import _compile, render_reference_kernel   # noqa: E402


class _Part:
    """One grid: model library, state, idiff, logged intermediaries."""

    def __init__(self, model, precision, kernel, dims, paced, g, inter_log,
                 fiber):
        self.nx, self.ny = dims
        self.n = self.nx * self.ny
        (self.model, self.vm, _rl, self.bound) = cgen.prepare_model(
            model, True, False)
        self.nstate = self.model.count_states()
        self.ivm = self.vm.index()
        self.gx, self.gy = float(g[0]), float(g[1])
        self.inter_log = [self.model.get(q) for q in inter_log]
        if kernel == 'ref':
            # fiber_tissue.py:867-884: explicit paced list / no paced cells
            text = render_reference_kernel(
                self.model, precision, self.bound, self.inter_log, True, [],
                list(paced), {}, False, False, fiber_tissue=fiber)
        else:
            text = cgen.generate(self.model, precision, self.bound,
                                 self.inter_log, True, [], {})
        so = _compile(text, 'oracle_ft_%s_%d' % (kernel, precision),
                      kernel == 'ref', False, False, '-O2')
        self.lib = ctypes.CDLL(so)
        self.real = np.float32 if self.lib.oracle_real_size() == 4 else np.float64
        self.state = np.tile(np.array(
            self.model.initial_values(True), dtype=self.real), self.n)
        self.idiff = np.zeros(self.n + 1, dtype=self.real)
        self.inter = np.zeros(self.n * max(len(self.inter_log), 1) + 1,
                              dtype=self.real)
        self.field = np.zeros(1, dtype=self.real)
        self.mask = np.zeros(self.n, dtype=np.uint8)
        for cid in paced:
            self.mask[cid] = 1

    def p(self, a):
        return a.ctypes.data_as(ctypes.c_void_p)

    def diff(self):
        self.lib.oracle_k_diff(
            ctypes.c_size_t(self.nx), ctypes.c_size_t(self.ny),
            ctypes.c_double(self.gx), ctypes.c_double(self.gy),
            self.p(self.state), self.p(self.idiff))

    def cells(self, time, dt, pace):
        self.lib.oracle_k_cells(
            ctypes.c_size_t(self.nx), ctypes.c_size_t(self.ny),
            ctypes.c_double(time), ctypes.c_double(dt), ctypes.c_double(pace),
            self.p(self.mask), self.p(self.state), self.p(self.idiff),
            self.p(self.inter), self.p(self.field))


class OracleFiberTissue:
    """
    ``run(duration, logf, logt, log_interval)`` returns
    ``(time, fiber_fields, tissue_fields)``: ``time`` (nt,), each ``*_fields``
    a dict ``qname -> (nt, ny, nx)`` array for the requested states, logged
    intermediaries or the diffusion-current variable.
    """

    def __init__(self, fiber_model, tissue_model, protocol=None,
                 ncells_fiber=(128, 2), ncells_tissue=(128, 128), nx_paced=5,
                 g_fiber=(9, 6), g_tissue=(9, 6), g_fiber_tissue=9, dt=0.005,
                 precision=myokit.SINGLE_PRECISION, kernel='port',
                 inter_log_fiber=(), inter_log_tissue=()):
        nfx, nfy = [int(x) for x in ncells_fiber]
        ntx, nty = [int(x) for x in ncells_tissue]
        if nfy > nty:
            raise ValueError(
                'The fiber y-dimension cannot exceed that of the tissue.')
        # fiber_tissue.py:191-196
        nx_paced = min(int(nx_paced), nfx)
        paced = [x + y * nfx for y in range(nfy) for x in range(nx_paced)]
        self._f = _Part(fiber_model, precision, kernel, (nfx, nfy), paced,
                        g_fiber, inter_log_fiber, True)
        self._t = _Part(tissue_model, precision, kernel, (ntx, nty), [],
                        g_tissue, inter_log_tissue, False)
        if self._f.real is not self._t.real:
            raise RuntimeError('precision mismatch')
        self._gft = float(g_fiber_tissue)
        # fiber_tissue.py:219-222
        self._ctx = 0
        self._cty = int(0.5 * (nty - nfy))
        self._dt = float(dt)
        self._time = 0.0
        self._protocol = None if protocol is None else protocol.clone()
        self.last_steps = 0

    def fiber_state(self):
        return np.array(self._f.state, dtype=np.float64)

    def tissue_state(self):
        return np.array(self._t.state, dtype=np.float64)

    def set_fiber_state(self, state):
        self._f.state[:] = np.asarray(state, dtype=np.float64).ravel()

    def set_tissue_state(self, state):
        self._t.state[:] = np.asarray(state, dtype=np.float64).ravel()

    def time(self):
        return self._time

    def set_time(self, time=0):
        self._time = float(time)

    def _events(self):
        ev = []
        if self._protocol is not None:
            for e in self._protocol.events():
                ev.extend([e.level(), e.start(), e.duration(), e.period(),
                           e.multiplier()])
        return np.array(ev if ev else [0.0] * 5, dtype=np.float64), len(ev) // 5

    def _junction(self):
        f, t = self._f, self._t
        f.lib.oracle_k_junction(
            ctypes.c_size_t(f.nx), ctypes.c_size_t(f.ny), ctypes.c_size_t(t.nx),
            ctypes.c_size_t(self._ctx), ctypes.c_size_t(self._cty),
            ctypes.c_size_t(t.nstate), ctypes.c_int(t.ivm),
            ctypes.c_double(self._gft), f.p(f.state), t.p(t.state),
            f.p(f.idiff), t.p(t.idiff))

    @staticmethod
    def _columns(part, names):
        """(kind, index) for each requested variable of one part."""
        vdiff = part.model.binding('diffusion_current')
        out = []
        for name in names:
            var = part.model.get(name)
            if var is vdiff:
                out.append(('idiff', 0))
            elif var.is_state():
                out.append(('state', var.index()))
            elif var in part.inter_log:
                out.append(('inter', part.inter_log.index(var)))
            else:
                raise ValueError('cannot log ' + name)
        return out

    def run(self, duration, logf=(), logt=(), log_interval=1.0):
        f, t = self._f, self._t
        real = f.real
        tmin = self._time
        tmax = tmin + duration
        default_dt = self._dt
        log_interval = 1e-9 if not log_interval or log_interval <= 0 \
            else float(log_interval)
        colsf = self._columns(f, logf)
        colst = self._columns(t, logt)
        rows_t, rows_f, rows_time = [], [], []
        events, n_events = self._events()
        rc = ctypes.c_int(0)
        f.lib.oracle_pacing_new.restype = ctypes.c_void_p
        pacing = ctypes.c_void_p(f.lib.oracle_pacing_new(
            ctypes.c_double(tmin), ctypes.c_int(n_events),
            events.ctypes.data_as(ctypes.c_void_p), ctypes.byref(rc)))
        if rc.value:
            raise RuntimeError('Oracle pacing error %d' % rc.value)
        level = ctypes.c_double(0)
        tnext = ctypes.c_double(0)
        f.lib.oracle_pacing_state(pacing, ctypes.byref(level), ctypes.byref(tnext))
        # fiber_tissue.c:473-496
        engine_time = tmin
        engine_pace = level.value
        tnext_pace = tnext.value
        arg_time = float(real(engine_time))
        arg_pace = float(real(engine_pace))
        istep = 1
        inext_log = 0
        tnext_log = tmin
        dt_min = 0.0
        halt = False
        steps = 0

        def grab(part, cols, snap):
            out = []
            for kind, k in cols:
                if kind == 'state':
                    a = snap.reshape(part.n, part.nstate)[:, k]
                elif kind == 'idiff':
                    a = part.idiff[:part.n]
                else:
                    a = part.inter[:part.n * len(part.inter_log)].reshape(
                        part.n, len(part.inter_log))[:, k]
                out.append(np.array(a, dtype=np.float64).reshape(part.ny, part.nx))
            return out

        while duration > 0:
            # fiber_tissue.c:1016-1027
            logging = engine_time >= tnext_log
            intermediary = False
            dt = tmin + float(istep) * default_dt - engine_time
            for d in (tmax - engine_time, tnext_pace - engine_time,
                      tnext_log - engine_time):
                if d > dt_min and d < dt:
                    dt = d
                    intermediary = True
            if not intermediary:
                istep += 1
            # :1031-1033
            f.diff()
            t.diff()
            self._junction()
            # :1036-1052
            snapf = snapt = None
            if logging:
                snapf = f.state.copy()
                snapt = t.state.copy()
                if np.isnan(snapf[0]) or np.isnan(snapt[0]):
                    halt = True
            # :1055-1062 (time, dt, pace are cast to Real)
            f.cells(arg_time, float(real(dt)), arg_pace)
            t.cells(arg_time, float(real(dt)), arg_pace)
            steps += 1
            # :1065-1119
            if logging:
                rows_time.append(arg_time)
                rows_f.append(grab(f, colsf, snapf))
                rows_t.append(grab(t, colst, snapt))
                inext_log += 1
                tnext_log = tmin + float(inext_log) * log_interval
            # :1130-1140
            engine_time += dt
            arg_time = float(real(engine_time))
            rc = f.lib.oracle_pacing_advance(
                pacing, ctypes.c_double(engine_time), ctypes.byref(level),
                ctypes.byref(tnext))
            if rc:
                f.lib.oracle_pacing_free(pacing)
                raise RuntimeError('Oracle pacing error %d' % rc)
            tnext_pace = tnext.value
            engine_pace = level.value
            arg_pace = float(real(engine_pace))
            if engine_time >= tmax or halt:
                break
        f.lib.oracle_pacing_free(pacing)
        self._time = tmax
        self.last_steps = steps
        nt = len(rows_time)

        def pack(part, names, rows):
            return dict(
                (name, np.array([r[k] for r in rows]).reshape(nt, part.ny, part.nx))
                for k, name in enumerate(names))
        return (np.array(rows_time), pack(f, list(logf), rows_f),
                pack(t, list(logt), rows_t))
