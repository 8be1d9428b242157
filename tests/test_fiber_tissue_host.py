"""
Host-side behaviour of FiberTissueSimulationCUDA (no GPU): argument checks with
the reference's messages (myokit/_sim/fiber_tissue.py:119-325, asserted by the
reference's tests/test_simulation_fiber_tissue.py::test_creation), state and
time handling, the kernels it would compile, and the loud failure without a
device.
"""
import os

import numpy as np
import pytest

import myokit_b200
import myokit
from myokit_b200 import capi

DATA = os.path.join(os.path.dirname(myokit.__file__), 'tests', 'data')
FT = myokit_b200.FiberTissueSimulationCUDA


def models():
    mf = myokit.load_model(os.path.join(DATA, 'dn-1985-normalised.mmt'))
    mt = myokit.load_model(os.path.join(DATA, 'lr-1991.mmt'))
    return mf, mt


def make(**kw):
    mf, mt = models()
    p = myokit.pacing.blocktrain(1000, 2.0, offset=.01)
    args = dict(ncells_fiber=(8, 4), ncells_tissue=(8, 6), nx_paced=4,
                g_fiber=(235, 100), g_tissue=(9, 5), g_fiber_tissue=9,
                dt=0.0012, precision=myokit.DOUBLE_PRECISION)
    args.update(kw)
    return FT(mf, mt, p, **args)


def test_creation_checks():
    mf, mt = models()
    s = make()
    assert s.fiber_shape() == (4, 8) and s.tissue_shape() == (6, 8)
    assert s.time() == 0 and s.step_size() == 0.0012
    with pytest.raises(ValueError, match='fiber size must be a tuple'):
        make(ncells_fiber=4)
    with pytest.raises(ValueError, match='tissue size must be a tuple'):
        make(ncells_tissue=(4, 4, 4))
    with pytest.raises(ValueError, match='fiber size must be at least'):
        make(ncells_fiber=(0, 4))
    with pytest.raises(ValueError, match='tissue size must be at least'):
        make(ncells_tissue=(8, 0))
    with pytest.raises(ValueError, match='fiber y-dimension cannot exceed'):
        make(ncells_fiber=(8, 7))
    with pytest.raises(ValueError, match='stimulus pulse must be non-negative'):
        make(nx_paced=-1)
    with pytest.raises(ValueError, match='fiber conductivity must be a tuple'):
        make(g_fiber=1)
    with pytest.raises(ValueError, match='tissue conductivity must be a tuple'):
        make(g_tissue=(1, 2, 3))
    with pytest.raises(ValueError, match='step size must be greater than zero'):
        make(dt=0)
    with pytest.raises(ValueError, match='single and double precision'):
        make(precision=16)
    # labels, bindings, units
    m2 = mt.clone()
    m2.label('membrane_potential').set_label(None)
    with pytest.raises(ValueError, match='labelled as "membrane_potential" in the fiber'):
        FT(m2, mt)
    with pytest.raises(ValueError, match='labelled as "membrane_potential" in the tissue'):
        FT(mt, m2)
    m2 = mt.clone()
    m2.binding('diffusion_current').set_binding(None)
    with pytest.raises(ValueError, match='bound to "diffusion_current"'):
        FT(mt, m2)
    m2 = mt.clone()
    m2.label('membrane_potential').set_unit(None)
    with pytest.raises(ValueError, match='fiber model must specify a unit for the membrane'):
        FT(m2, mt)
    m2 = mt.clone()
    m2.label('membrane_potential').set_unit('V')
    with pytest.raises(ValueError, match='same unit in the fiber and the tissue'):
        FT(mt, m2)


def test_states_time_and_reset():
    s = make()
    mf, mt = models()
    nf, nt = mf.count_states(), mt.count_states()
    assert len(s.fiber_state()) == nf * 32 and len(s.tissue_state()) == nt * 48
    assert list(s.fiber_state(1, 2)) == list(mf.initial_values(True))
    st = list(mt.initial_values(True))
    st[0] = -50.0
    s.set_tissue_state(st, 3, 4)
    assert s.tissue_state(3, 4)[0] == -50.0
    assert s.tissue_state(2, 4)[0] != -50.0
    assert s.default_tissue_state(3, 4)[0] != -50.0
    s.set_default_fiber_state(s.fiber_state())
    s.set_time(12.5)
    assert s.time() == 12.5
    s.reset()
    assert s.time() == 0 and s.tissue_state(3, 4)[0] != -50.0
    s.set_step_size(0.002)
    assert s.step_size() == 0.002
    with pytest.raises(ValueError):
        s.set_step_size(0)
    with pytest.raises(ValueError, match="can't be negative"):
        s.run(-1)


def test_kernels_carry_their_side_of_the_junction():
    s = make()
    cf = s._f.kernel_source().code
    ct = s._t.kernel_source().code
    assert 'idiff += (Real)g.jg * (vc - vo);' in cf
    assert 'idiff -= (Real)g.jg * (vo - vc);' in ct
    # one cell per thread even for tiny models, fibre paced over its height
    assert '#define MKB_CPT' not in cf and '#define MKB_CPT' not in ct
    assert s._f._paced_cells == (4, 4, 0, 0) and s._t._paced_cells == (0, 0, 0, 0)
    for src in (s._f.kernel_source(), s._t.kernel_source()):
        cubin, log = capi.jit_compile(src.code, src.options)
        assert len(cubin) > 5000


def test_no_device_no_run():
    if capi.device_count() > 0:
        pytest.skip('a GPU is present')
    s = make()
    with pytest.raises(Exception):
        s.run(1.0)
    # nothing was advanced or left half-open
    assert s.time() == 0 and s._f._session is None and s._t._session is None


class _RecordingBackend:
    """
    Stands in for the device entry points so that the host-side control flow of
    a pair run (acquire both, connect once, step until done, read both logs,
    re-arm on the next run) can be exercised without a GPU. Everything else
    (JIT, schedule probe, ...) goes to the real library. It computes nothing.
    """

    def __init__(self, real):
        self._real = real
        self.calls = []
        self._next = 0x1000
        self._log = {}

    def __getattr__(self, name):
        return getattr(self._real, name)

    def mkb_sim_init(self, cfg, out):
        import ctypes
        c = ctypes.cast(cfg, ctypes.POINTER(capi.SimConfig)).contents
        self._next += 0x100
        ctypes.cast(out, ctypes.POINTER(ctypes.c_void_p)).contents.value = self._next
        self.calls.append(('init', self._next, c.nx, c.ny, c.n_log))
        self._log[self._next] = (c.n_log, c.tmin, c.tmax, c.log_interval)
        return 0

    def mkb_sim_rearm(self, sim, rc):
        import ctypes
        r = ctypes.cast(rc, ctypes.POINTER(capi.RunConfig)).contents
        self.calls.append(('rearm', sim.value))
        self._log[sim.value] = (r.n_log, r.tmin, r.tmax, r.log_interval)
        return 0

    def mkb_sim_set_state(self, sim, state, uniform):
        self.calls.append(('set_state', sim.value, uniform))
        return 0

    def mkb_sim_junction_connect(self, f, t, g, cty):
        self.calls.append(('connect', f.value, t.value, g.value, cty.value))
        return 0

    def mkb_sim_step_pair(self, f, t, steps, now, halted):
        import ctypes
        self.calls.append(('step_pair', f.value, t.value, steps.value))
        done = sum(1 for c in self.calls if c[0] == 'step_pair'
                   and c[1] == f.value) % 3 == 0
        tmin, tmax = self._log[f.value][1:3]
        ctypes.cast(now, ctypes.POINTER(ctypes.c_double)).contents.value = \
            tmax if done else 0.5 * (tmin + tmax)
        return 0 if done else 1

    def mkb_sim_log_view(self, sim, data, rows, cols, rstride):
        import ctypes
        import numpy as np
        n_log, tmin, tmax, li = self._log[sim.value]
        nrows = int(round((tmax - tmin) / li))
        buf = np.zeros((nrows, n_log + 1))
        buf[:, 0] = tmin + li * np.arange(nrows)
        self._keep = getattr(self, '_keep', []) + [buf]
        ctypes.cast(data, ctypes.POINTER(ctypes.c_void_p)).contents.value = buf.ctypes.data
        ctypes.cast(rows, ctypes.POINTER(ctypes.c_uint64)).contents.value = nrows
        ctypes.cast(cols, ctypes.POINTER(ctypes.c_uint64)).contents.value = n_log
        ctypes.cast(rstride, ctypes.POINTER(ctypes.c_uint64)).contents.value = n_log + 1
        return 0

    def mkb_sim_counters(self, sim, launches, steps):
        return 0

    def mkb_sim_device_ms(self, sim, ms):
        return 0

    def mkb_sim_get_state(self, sim, out):
        self.calls.append(('get_state', sim.value))
        return 0

    def mkb_sim_clean(self, sim):
        self.calls.append(('clean', sim.value if hasattr(sim, 'value') else sim))


def test_pair_run_control_flow(monkeypatch):
    real = capi.library()
    fake = _RecordingBackend(real)
    monkeypatch.setattr(capi, 'library', lambda: fake)
    s = make()
    logf, logt = s.run(2.0, logf=['engine.time', 'membrane.V'],
                       logt=['engine.time', 'membrane.V'], log_interval=0.5)
    kinds = [c[0] for c in fake.calls]
    assert kinds[:3] == ['init', 'init', 'connect']
    assert fake.calls[0][2:4] == (8, 4) and fake.calls[1][2:4] == (8, 6)
    f, t = fake.calls[0][1], fake.calls[1][1]
    assert fake.calls[2] == ('connect', f, t, 9.0, 1)        # cty = (6 - 4) / 2
    assert kinds[3:] == ['step_pair'] * 3
    assert all(c[1:3] == (f, t) for c in fake.calls[3:])
    # both logs: 4 rows, every cell of the requested variable
    assert list(logf['engine.time']) == [0.0, 0.5, 1.0, 1.5]
    assert len(logf['7.3.membrane.V']) == 4 and len(logt['7.5.membrane.V']) == 4
    assert len(logf.keys()) == 1 + 32 and len(logt.keys()) == 1 + 48
    assert s.time() == 2.0 and s._f.time() == 2.0 and s._t.time() == 2.0
    # second run: both re-armed on the resident state, no second connect
    n0 = len(fake.calls)
    s.run(1.0, logf=['membrane.V'], logt=myokit.LOG_NONE, log_interval=0.5)
    kinds = [c[0] for c in fake.calls[n0:]]
    assert kinds[:2] == ['rearm', 'rearm'] and 'connect' not in kinds
    assert s.time() == 3.0
    # a new state goes into the resident pair: no restart, no new junction
    n0 = len(fake.calls)
    s.set_tissue_state(s.tissue_state())
    s.run(1.0, logf=myokit.LOG_NONE, logt=myokit.LOG_NONE)
    kinds = [c[0] for c in fake.calls[n0:]]
    assert kinds.count('set_state') == 1 and kinds.count('rearm') == 2
    assert 'init' not in kinds and 'connect' not in kinds
    # touching what the kernels were built for restarts both and joins the new pair
    n0 = len(fake.calls)
    s._t.set_conductance(3, 2)
    s.run(1.0, logf=myokit.LOG_NONE, logt=myokit.LOG_NONE)
    kinds = [c[0] for c in fake.calls[n0:]]
    assert kinds.count('clean') == 2 and kinds.count('init') == 2
    assert kinds.count('connect') == 1
    # pre: time stays, defaults follow
    s.pre(1.0)
    assert s.time() == 5.0
    s.close()


def test_find_nan_on_logs():
    s = make()
    mf, mt = models()
    nt_ = 5

    def full_log(model, nx, ny):
        log = myokit.DataLog()
        log['engine.time'] = list(np.arange(nt_) * 0.5)
        log['engine.pace'] = [0.0] * nt_
        for y in range(ny):
            for x in range(nx):
                pre = '%d.%d.' % (x, y)
                for st in model.states():
                    log[pre + st.qname()] = [0.1 * k for k in range(nt_)]
                log[pre + model.binding('diffusion_current').qname()] = [0.0] * nt_
        return log
    logf = full_log(s._f._model, 8, 4)
    logt = full_log(s._t._model, 8, 6)
    with pytest.raises(myokit.FindNanError, match='not found'):
        s.find_nan(logf, logt)
    vt = s._t._model.label('membrane_potential').qname()
    logt['3.2.' + vt][3] = float('nan')
    logt['3.2.' + vt][4] = float('nan')
    logf['1.1.' + s._f._model.label('membrane_potential').qname()][4] = float('inf')
    part, time, icell, var, value, states, bound = s.find_nan(logf, logt)
    assert part == 'tissue' and time == 1.5 and icell == (3, 2) and var == vt
    assert value != value and len(states) == 3 and len(bound) == 3
    assert len(states[0]) == s._t._model.count_states()
    assert bound[0]['engine.time'] == 1.5 and bound[1]['engine.time'] == 1.0
    del logt['0.0.' + vt]
    with pytest.raises(myokit.FindNanError, match='tissue model containing all states'):
        s.find_nan(logf, logt)
