"""
``FiberTissueSimulationCUDA``: the B200 counterpart of
``myokit.FiberTissueSimulation`` (``myokit/_sim/fiber_tissue.py:17``).

A 2-d fibre and a 2-d tissue, each with its own model, take every time step
together; the last fibre column drives tissue cells ``(0, cty + y)`` through
``g_fiber_tissue`` (``myokit/_sim/openclsim.cl:601-628``). Here the two grids
are two :class:`SimulationCUDA` back-ends whose kernels carry the junction
term (``kernelgen.generate(junction=...)``), connected and stepped through
``mkb_sim_junction_connect`` / ``mkb_sim_step_pair``.

Status: the generated junction kernels reproduce the fibre-tissue oracle bit
for bit when run on the host (``tests/test_generated_kernel_host.py``); the
device path of this class has not run on a GPU yet (its GPU tests are marked
experimental).
"""
import ctypes

import numpy as np

import myokit

from . import capi
from .simulation import SimulationCUDA


class FiberTissueSimulationCUDA:
    """
    See :class:`myokit.FiberTissueSimulation` for the arguments; error
    messages follow ``fiber_tissue.py:119-325``. Forward Euler, homogeneous
    conductances, one GPU.
    """

    def __init__(self, fiber_model, tissue_model, protocol=None,
                 ncells_fiber=(128, 2), ncells_tissue=(128, 128), nx_paced=5,
                 g_fiber=(9, 6), g_tissue=(9, 6), g_fiber_tissue=9,
                 dt=0.005, precision=myokit.SINGLE_PRECISION,
                 native_maths=False, device=0):
        fiber_model.validate()
        tissue_model.validate()
        for model, name in ((fiber_model, 'fiber'), (tissue_model, 'tissue')):
            if model.has_interdependent_components():
                cycles = '\n'.join([
                    '  ' + ' > '.join([x.name() for x in c])
                    for c in model.component_cycles()])
                raise ValueError(
                    'This simulation requires models without interdependent'
                    ' components. Please restructure the ' + name + ' model'
                    ' and re-run. Cycles:\n' + cycles)

        for cells, name in ((ncells_fiber, 'fiber'), (ncells_tissue, 'tissue')):
            msg = 'The ' + name + ' size must be a tuple (nx, ny).'
            try:
                if len(cells) != 2:
                    raise ValueError(msg)
            except TypeError:
                raise ValueError(msg)
        self._ncellsf = [int(x) for x in ncells_fiber]
        self._ncellst = [int(x) for x in ncells_tissue]
        if self._ncellsf[0] < 1 or self._ncellsf[1] < 1:
            raise ValueError('The fiber size must be at least (1, 1).')
        if self._ncellst[0] < 1 or self._ncellst[1] < 1:
            raise ValueError('The tissue size must be at least (1, 1).')
        if self._ncellsf[1] > self._ncellst[1]:
            raise ValueError(
                'The fiber y-dimension cannot exceed that of the tissue.')

        nx_paced = int(nx_paced)
        if nx_paced < 0:
            raise ValueError(
                'The width of the stimulus pulse must be non-negative.')
        nx_paced = min(nx_paced, self._ncellsf[0])

        for g, name in ((g_fiber, 'fiber'), (g_tissue, 'tissue')):
            msg = 'The ' + name + ' conductivity must be a tuple (gx, gy).'
            try:
                if len(g) != 2:
                    raise ValueError(msg)
            except TypeError:
                raise ValueError(msg)
        self._gf = [float(x) for x in g_fiber]
        self._gt = [float(x) for x in g_tissue]
        self._gft = float(g_fiber_tissue)

        # Point of connection to the tissue (fiber_tissue.py:219-222)
        self._cfx = self._ncellsf[0] - 1
        self._ctx = 0
        self._cty = int(0.5 * (self._ncellst[1] - self._ncellsf[1]))

        dt = float(dt)
        if dt <= 0:
            raise ValueError('The step size must be greater than zero.')
        if precision not in (myokit.SINGLE_PRECISION, myokit.DOUBLE_PRECISION):
            raise ValueError('Only single and double precision are supported.')
        self._precision = precision

        # Labels, bindings and units the two models must agree on
        # (fiber_tissue.py:243-296)
        for model, name in ((fiber_model, 'fiber'), (tissue_model, 'tissue')):
            vm = model.label('membrane_potential')
            if vm is None:
                raise ValueError(
                    'This simulation requires the membrane potential variable'
                    ' to be labelled as "membrane_potential" in the ' + name
                    + ' model.')
            if not vm.is_state():
                raise ValueError(
                    'The variable labelled as membrane potential in the '
                    + name + ' model must be a state variable.')
            if model.binding('diffusion_current') is None:
                raise ValueError(
                    'This simulation requires a variable in the ' + name
                    + ' model to be bound to "diffusion_current" to pass'
                    ' current from one cell to the next.')
        uvf = fiber_model.label('membrane_potential').unit()
        uvt = tissue_model.label('membrane_potential').unit()
        if uvf is None:
            raise ValueError('The fiber model must specify a unit for the'
                             ' membrane potential.')
        if uvt is None:
            raise ValueError('The tissue model must specify a unit for the'
                             ' membrane potential.')
        if uvf != uvt:
            raise ValueError(
                'The membrane potential must have the same unit in the fiber'
                ' and the tissue model: ' + str(uvf) + ' vs ' + str(uvt) + '.')
        ucf = fiber_model.binding('diffusion_current').unit()
        uct = tissue_model.binding('diffusion_current').unit()
        if ucf is None:
            raise ValueError('The fiber model must specify a unit for the'
                             ' diffusion current.')
        if uct is None:
            raise ValueError('The tissue model must specify a unit for the'
                             ' diffusion current.')
        if ucf != uct:
            raise ValueError(
                'The diffusion current must have the same unit in the fiber'
                ' and the tissue model: ' + str(ucf) + ' vs ' + str(uct) + '.')

        # The two grids
        self._f = SimulationCUDA(
            fiber_model, protocol, ncells=tuple(self._ncellsf),
            precision=precision, native_maths=native_maths, device=device)
        self._t = SimulationCUDA(
            tissue_model, protocol, ncells=tuple(self._ncellst),
            precision=precision, native_maths=native_maths, device=device)
        self._f.set_conductance(*self._gf)
        self._t.set_conductance(*self._gt)
        # the stimulus covers the first nx_paced columns of the fibre, full
        # height (fiber_tissue.py:191-196); the tissue is never paced (:880)
        self._f.set_paced_cells(nx_paced, self._ncellsf[1], 0, 0)
        self._t.set_paced_cells(0, 0, 0, 0)
        self._f.set_kernel_options(junction='fiber')
        self._t.set_kernel_options(junction='tissue')
        self._time = 0
        self.set_step_size(dt)
        self._connected = None      # the pair of back-end handles last joined

    # ------------------------------------------------------------------
    # State, time, protocol
    # ------------------------------------------------------------------
    def default_fiber_state(self, x=None, y=None):
        return self._f.default_state(x, y)

    def default_tissue_state(self, x=None, y=None):
        return self._t.default_state(x, y)

    def fiber_shape(self):
        """Shape of the fibre as ``(ny, nx)`` (fiber_tissue.py:433-439)."""
        return (self._ncellsf[1], self._ncellsf[0])

    def tissue_shape(self):
        """Shape of the tissue as ``(ny, nx)``."""
        return (self._ncellst[1], self._ncellst[0])

    def fiber_state(self, x=None, y=None):
        return self._f.state(x, y)

    def tissue_state(self, x=None, y=None):
        return self._t.state(x, y)

    def set_fiber_state(self, state, x=None, y=None):
        self._f.set_state(state, x, y)

    def set_tissue_state(self, state, x=None, y=None):
        self._t.set_state(state, x, y)

    def set_default_fiber_state(self, state, x=None, y=None):
        self._f.set_default_state(state, x, y)

    def set_default_tissue_state(self, state, x=None, y=None):
        self._t.set_default_state(state, x, y)

    def set_step_size(self, step_size=0.005):
        step_size = float(step_size)
        if step_size <= 0:
            raise ValueError('Step size must be greater than zero.')
        self._f.set_step_size(step_size)
        self._t.set_step_size(step_size)
        self._step_size = step_size

    def step_size(self):
        return self._step_size

    def set_protocol(self, protocol=None):
        self._f.set_protocol(protocol)
        self._t.set_protocol(protocol)

    def set_time(self, time=0):
        self._time = float(time)
        self._f.set_time(self._time)
        self._t.set_time(self._time)

    def time(self):
        return self._time

    def reset(self):
        self._f.reset()
        self._t.reset()
        self._time = 0

    def close(self):
        """Releases the device memory of both grids (states are kept)."""
        self._f.close()
        self._t.close()
        self._connected = None

    # ------------------------------------------------------------------
    # NaN search
    # ------------------------------------------------------------------
    def find_nan(self, logf, logt):
        """
        Locates the first ``NaN`` / ``inf`` in a pair of logs made by this
        simulation (cf. ``fiber_tissue.py:464-687``). The logs must hold all
        states and bound variables of both parts.

        Returns ``(part, time, icell, variable, value, states, bound)``:
        ``part`` is ``'fiber'`` or ``'tissue'``, ``icell`` an ``(x, y)``
        tuple, ``states[0]`` the state of that cell at the logged point where
        the error shows, ``states[1]``, ``states[2]`` the logged points before
        it (if any), ``bound`` the bound variables at the same points.

        Unlike the reference, the search works on the logged samples only: it
        does not re-run the preceding interval with a finer log.
        """
        parts = (('fiber', self._f, logf), ('tissue', self._t, logt))
        for name, sim, log in parts:
            g = [v.qname() for v in (sim._model.binding(x) for x in
                                     ('time', 'pace')) if v is not None]
            need = myokit.prepare_log(
                myokit.LOG_STATE + myokit.LOG_BOUND, sim._model,
                dims=sim._dims, global_vars=g)
            for key in need:
                if key not in log:
                    raise myokit.FindNanError(
                        'Method requires a simulation log from the ' + name
                        + ' model containing all states and bound variables.'
                        ' Missing variable <' + key + '>.')
        # first bad sample of every key, then the earliest over both logs
        best = None
        for name, sim, log in parts:
            for key in log.keys():
                bad = np.nonzero(~np.isfinite(np.asarray(log[key])))[0]
                if len(bad) and (best is None or bad[0] < best[0]):
                    best = (int(bad[0]), name, sim, log, key)
        if best is None:
            raise myokit.FindNanError('Error condition not found in logs.')
        i, name, sim, log, key = best
        if i == 0:
            raise myokit.FindNanError(
                'Unable to work with simulation logs where the error'
                ' condition is met in the very first data point.')
        cell, var = myokit.split_key(key)
        icell = tuple(int(x) for x in cell.split('.') if x != '')
        model = sim._model
        time_var = model.time().qname()
        prefix = '.'.join(str(x) for x in icell) + '.' if icell else ''
        if not icell:       # a global variable went wrong: report cell (0, 0)
            icell = (0, 0)
            prefix = '0.0.'
        states, bound = [], []
        for k in range(i, max(i - 3, -1), -1):
            states.append([log[prefix + s.qname()][k] for s in model.states()])
            b = {}
            for label in ('time', 'pace', 'diffusion_current'):
                v = model.binding(label)
                if v is None:
                    continue
                q = v.qname()
                b[q] = log[q][k] if q in log else log[prefix + q][k]
            bound.append(b)
        return (name, log[time_var][i], icell, var, log[key][i], states, bound)

    # ------------------------------------------------------------------
    # Running
    # ------------------------------------------------------------------
    def pre(self, duration, report_nan=True, progress=None,
            msg='Pre-pacing FiberTissueSimulationCUDA'):
        """
        Unlogged run that leaves the time unchanged and makes the final
        states the new default states (fiber_tissue.py:689-712).
        """
        self._run(duration, myokit.LOG_NONE, myokit.LOG_NONE, 1, report_nan,
                  progress, msg)
        for s in (self._f, self._t):
            s._sync_state()
            s._default_state = s._state.copy()

    def run(self, duration, logf=None, logt=None, log_interval=1.0,
            report_nan=True, progress=None,
            msg='Running FiberTissueSimulationCUDA'):
        """
        Runs for ``duration`` and returns ``(logf, logt)``: one
        :class:`myokit.DataLog` for the fibre and one for the tissue, keyed
        ``x.y.qname`` like ``myokit.FiberTissueSimulation.run``
        (fiber_tissue.py:727-784).
        """
        r = self._run(duration, logf, logt, log_interval, report_nan,
                      progress, msg)
        self._time += duration
        self._f._time = self._t._time = self._time
        return r

    def _prepare(self, sim, log):
        g = []
        for label in ('time', 'pace'):
            v = sim._model.binding(label)
            if v is not None:
                g.append(v.qname())
        log = myokit.prepare_log(
            log, sim._model, dims=sim._dims, global_vars=g,
            if_empty=myokit.LOG_STATE + myokit.LOG_BOUND,
            allowed_classes=myokit.LOG_STATE + myokit.LOG_BOUND
            + myokit.LOG_INTER, precision=self._precision)
        inter_log = []
        seen = set()
        for key in log.keys():
            name = myokit.split_key(key)[1]
            if name in seen:
                continue
            seen.add(name)
            var = sim._model.get(name)
            if var.is_intermediary() and not var.is_bound():
                inter_log.append(var)
        return log, inter_log

    def _run(self, duration, logf, logt, log_interval, report_nan, progress,
             msg):
        if duration < 0:
            raise ValueError('Simulation duration can\'t be negative.')
        tmin = self._time
        tmax = tmin + duration
        f, t = self._f, self._t
        f._time = t._time = tmin
        logf, inter_f = self._prepare(f, logf)
        logt, inter_t = self._prepare(t, logt)
        log_interval = 1e-9 if log_interval is None else float(log_interval)
        if log_interval <= 0:
            log_interval = 1e-9
        if progress is None:
            progress = myokit._simulation_progress
        if progress and not isinstance(progress, myokit.ProgressReporter):
            raise ValueError(
                'The argument "progress" must be either a subclass of'
                ' myokit.ProgressReporter or None.')
        halted = False
        if duration > 0:
            halted = self._run_pair(
                tmin, tmax, (logf, inter_f), (logt, inter_t), log_interval,
                progress, msg)
        if report_nan and (halted or logf.has_nan() or logt.has_nan()):
            txt = ['Numerical error found in simulation logs.']
            try:
                part, time, icell, var, value, states, bound = self.find_nan(
                    logf, logt)
                model = self._t._model if part == 'tissue' else self._f._model
                txt.append(
                    'Encountered numerical error in ' + part + ' simulation at'
                    ' t = ' + myokit.float.str(time, precision=self._precision)
                    + ' in cell (' + ','.join(str(x) for x in icell)
                    + ') when ' + var + ' = '
                    + myokit.float.str(value, precision=self._precision) + '.')
                txt.append('State during:')
                txt.append(model.format_state(
                    states[0], precision=self._precision))
                if len(states) > 1:
                    txt.append('State at the logged point before:')
                    txt.append(model.format_state(
                        states[1], precision=self._precision))
            except myokit.FindNanError as e:
                txt.append('Unable to pinpoint source of NaN, an error'
                           ' occurred:')
                txt.append(str(e))
            raise myokit.SimulationError('\n'.join(txt))
        return logf, logt

    def _run_pair(self, tmin, tmax, fiber, tissue, log_interval, progress,
                  msg):
        lib = capi.library()
        sims = (self._f, self._t)
        logs = (fiber[0], tissue[0])
        inters = (fiber[1], tissue[1])
        # A junction ties two particular back-ends together: unless both are
        # still resident (and will merely be re-armed), both start afresh
        for s in sims:
            if s._session is None or s._dirty:
                for q in sims:
                    q._close_session(fetch_state=True)
                    q._dirty = True
                self._connected = None
                break
        tables = [s._log_table(log, inter)
                  for s, log, inter in zip(sims, logs, inters)]
        handles = []
        try:
            for s, table, inter in zip(sims, tables, inters):
                keys, kinds, index = table
                handles.append(s._acquire(
                    tmin, tmax, keys, kinds, index, inter, log_interval))
            pair = tuple(int(h.value) if hasattr(h, 'value') else int(h)
                         for h in handles)
            if self._connected != pair:
                capi.check(lib.mkb_sim_junction_connect(
                    handles[0], handles[1], ctypes.c_double(self._gft),
                    ctypes.c_uint64(self._cty)))
                self._connected = pair
            halted = ctypes.c_int(0)
            now = ctypes.c_double(tmin)
            # like fiber_tissue.c:1012-1013: steps between progress updates
            chunk = max(40, int(6e6 / (self._f._ntotal + self._t._ntotal)))

            def loop(update):
                while True:
                    rc = capi.check(lib.mkb_sim_step_pair(
                        handles[0], handles[1], ctypes.c_uint64(chunk),
                        ctypes.byref(now), ctypes.byref(halted)))
                    if update is not None and not update(now.value):
                        raise myokit.SimulationCancelledError()
                    if rc == 0:
                        break
            if progress:
                with progress.job(msg):
                    r = 1.0 / (tmax - tmin)
                    loop(lambda x: progress.update(min((x - tmin) * r, 1)))
            else:
                loop(None)
            for s, sim, table, log in zip(sims, handles, tables, logs):
                keys = table[0]
                mat = s._log_matrix(lib, sim, len(keys))
                if mat is not None:
                    try:
                        cols = np.ascontiguousarray(mat.block(0, len(keys)).T)
                        for i, key in enumerate(keys):
                            log[key].extend(cols[i].tolist())
                    finally:
                        mat.release()
                s._info = s._collect_info(lib, sim)
                s._state_stale = True
        except BaseException:
            for s in sims:
                s._close_session(fetch_state=False)
                s._dirty = True
            self._connected = None
            raise
        return bool(halted.value)
