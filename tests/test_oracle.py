"""
Pins the CPU oracle (oracle/) — CPU only, no GPU needed.

1. against golden vectors produced by the REFERENCE itself
   (myokit.Simulation1d, tests/golden/make_golden.py) with the reference's own
   cross-implementation tolerances
   (myokit/tests/test_simulation_opencl_vs_sim1d.py:118-136: 1e-17 for time
   and pace, 1e-13 for V, i_diff, Isi);
2. against the reference's own rendered OpenCL kernel compiled as C
   (kernel='ref', oracle/_ref) on every feature the 1-d CPU reference cannot
   reach: 2-d grids, fp32, Rush-Larsen, conductance fields, scalar fields,
   connections, paced lists, logged intermediaries;
3. the pacing restatement against myokit's PacingSystem and the known answers
   of myokit/tests/test_pacing_system_c.py;
4. the time-step / log schedule against the properties the reference tests
   (myokit/tests/test_simulation_log_interval.py:21-62).
"""
import os

import numpy as np
import pytest

from oracle.oracle import OracleSimulation, pacing_probe, myokit

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
DP = myokit.DOUBLE_PRECISION
SP = myokit.SINGLE_PRECISION


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + '.npz'))
    log = dict((k[4:], z[k]) for k in z.files if k.startswith('log:'))
    meta = dict((k[5:], z[k]) for k in z.files if k.startswith('meta:'))
    return log, z['state'], meta


def data_model(name):
    return myokit.load_model(os.path.join(
        os.path.dirname(myokit.__file__), 'tests', 'data', name))


def maxdiff(a, b, keys):
    w = 0.0
    for k in keys:
        assert len(a[k]) == len(b[k]), k
        w = max(w, float(np.max(np.abs(np.asarray(a[k]) - np.asarray(b[k])))))
    return w


# ---------------------------------------------------------------------------
# 1. Golden vectors from the reference's Simulation1d
# ---------------------------------------------------------------------------
def test_golden_lr91_c1():
    log, state, meta = load_golden('sim1d_lr91_c1')
    m, p, _ = myokit.load('example')
    o = OracleSimulation(m, p, ncells=128, precision=DP)
    o.set_conductance(10)
    o.set_paced_cells(5)
    o.set_step_size(0.005)
    lg, st = o.run(120, log=['engine.time', 'engine.pace', 'membrane.V'],
                   log_interval=1)
    assert len(lg['engine.time']) == 120
    assert maxdiff(lg, log, ['engine.time', 'engine.pace']) < 1e-17
    vkeys = [k for k in log if k.endswith('membrane.V')]
    assert len(vkeys) == 128
    # The wave must actually be there: paced cells fire, far cells follow
    assert log['0.membrane.V'].max() > 0
    assert log['127.membrane.V'].max() > 0
    assert maxdiff(lg, log, vkeys) < 1e-13
    assert np.max(np.abs(st - state)) < 1e-13


def test_golden_br77_reference_contract():
    # myokit/tests/test_simulation_opencl_vs_sim1d.py:28-136
    log, state, meta = load_golden('sim1d_br77')
    m = data_model('beeler-1977-model.mmt')
    p = myokit.pacing.blocktrain(duration=2, offset=1, period=1000)
    o = OracleSimulation(m, p, ncells=10, precision=DP)
    o.set_conductance(10)
    o.set_paced_cells(3)
    o.set_step_size(0.005)
    lg, st = o.run(15, log=['engine.time', 'engine.pace', 'membrane.V',
                            'membrane.i_diff', 'isi.Isi'], log_interval=0.5)
    assert maxdiff(lg, log, ['engine.time', 'engine.pace']) < 1e-17
    for cell in (0, 9):
        for var in ('membrane.V', 'membrane.i_diff', 'isi.Isi'):
            assert maxdiff(lg, log, ['%d.%s' % (cell, var)]) < 1e-13
    assert maxdiff(lg, log, list(log.keys())) < 1e-13


def test_golden_lr91_rush_larsen():
    log, state, meta = load_golden('sim1d_lr91_rl')
    m, p, _ = myokit.load('example')
    o = OracleSimulation(m, p, ncells=32, precision=DP, rl=True)
    o.set_conductance(10)
    o.set_paced_cells(5)
    o.set_step_size(0.01)
    lg, st = o.run(80, log=['engine.time', 'membrane.V', 'ina.m'],
                   log_interval=1)
    assert maxdiff(lg, log, list(log.keys())) < 1e-13
    assert np.max(np.abs(st - state)) < 1e-13


def test_golden_decker_rush_larsen():
    # the bench model (48 states), Rush-Larsen, against the reference's
    # Simulation1d run in the build container
    log, state, meta = load_golden('sim1d_decker_rl')
    import os
    m = myokit.load_model(os.path.join(
        os.path.dirname(myokit.__file__), 'tests', 'data', 'decker-2009.mmt'))
    p = myokit.pacing.blocktrain(duration=2, offset=1, period=1000)
    for kernel in ('port', 'ref'):
        o = OracleSimulation(m, p, ncells=12, precision=DP, rl=True,
                             kernel=kernel)
        o.set_conductance(10)
        o.set_paced_cells(3)
        o.set_step_size(0.005)
        lg, st = o.run(12, log=['engine.time', 'engine.pace', 'membrane.V',
                                'membrane.i_diff', 'ina.m', 'calcium.uCa_i'],
                       log_interval=0.5)
        assert np.max(log['0.membrane.V']) > 0       # the paced end fired
        assert maxdiff(lg, log, list(log.keys())) < 1e-12
        rel = np.abs(st - state) / (np.abs(state) + 1e-12)
        assert rel.max() < 1e-12


# ---------------------------------------------------------------------------
# 2. Port vs the reference's own rendered kernel (oracle/_ref)
# ---------------------------------------------------------------------------
def pair(model, protocol, ncells, duration, log, setup, precision=DP,
         diffusion=True, rl=False, log_interval=0.5):
    out = []
    for kernel in ('port', 'ref'):
        o = OracleSimulation(model, protocol, ncells=ncells,
                             precision=precision, diffusion=diffusion, rl=rl,
                             kernel=kernel)
        setup(o)
        out.append(o.run(duration, log=log, log_interval=log_interval))
    (la, sa), (lb, sb) = out
    assert set(la.keys()) == set(lb.keys())
    worst = maxdiff(la, lb, list(la.keys()))
    return worst, float(np.max(np.abs(sa - sb))), la


PULSE = dict(duration=2, offset=1, period=1000)


def test_ref_2d_fp64_homogeneous():
    m, _, _ = myokit.load('example')
    p = myokit.pacing.blocktrain(**PULSE)

    def setup(o):
        o.set_conductance(10, 4)
        o.set_paced_cells(3, 12, 0, 0)
    w, ws, la = pair(m, p, (20, 12), 12, ['engine.time', 'membrane.V',
                                          'membrane.i_diff', 'ica.ICa'], setup)
    assert la['19.11.membrane.V'].max() > 0   # wave crossed the grid
    assert w == 0 and ws == 0


def test_ref_2d_fp32_rl_paced_list():
    m, _, _ = myokit.load('example')
    p = myokit.pacing.blocktrain(**PULSE)

    def setup(o):
        o.set_conductance(8, 8)
        o.set_paced_cell_list([(0, 0), (1, 0), (0, 1), (1, 1), (2, 2)])
        o.set_step_size(0.01)
    w, ws, la = pair(m, p, (12, 10), 10, ['engine.time', 'engine.pace',
                                          'membrane.V'], setup,
                     precision=SP, rl=True)
    assert la['0.0.membrane.V'].max() > 0
    assert w == 0 and ws == 0


def test_ref_hetero_fields_decker():
    m = data_model('decker-2009.mmt')
    p = myokit.pacing.blocktrain(**PULSE)
    rng = np.random.default_rng(5)
    gx = rng.uniform(5, 12, size=(6, 9))
    gy = rng.uniform(2, 7, size=(5, 10))
    gx[2, 3:6] = 0
    gk = 0.0138542 * rng.uniform(0.5, 1.5, size=(6, 10))

    def setup(o):
        o.set_conductance_field(gx, gy)
        o.set_field('ikr.Gbar', gk)
        o.set_paced_cells(2, 6, 0, 0)
    w, ws, la = pair(m, p, (10, 6), 6, ['engine.time', 'membrane.V',
                                        'ikr.IKr'], setup, rl=True)
    assert la['9.5.membrane.V'].max() > 0
    assert w == 0 and ws == 0


def test_ref_connections_match_and_equal_conductance():
    m, _, _ = myokit.load('example')
    p = myokit.pacing.blocktrain(**PULSE)
    n = 12
    conns = [(i, i + 1, 9.0) for i in range(n - 1)]

    def setup(o):
        o.set_connections(conns)
        o.set_paced_cells(2)
    w, ws, la = pair(m, p, n, 8, ['engine.time', 'membrane.V',
                                  'membrane.i_diff'], setup)
    assert w == 0 and ws == 0
    # connections == conductance to 1e-9 (fp64), test_simulation_opencl.py:525-528
    o = OracleSimulation(m, p, ncells=n, precision=DP)
    o.set_conductance(9.0)
    o.set_paced_cells(2)
    lb, sb = o.run(8, log=['engine.time', 'membrane.V'], log_interval=0.5)
    vk = [k for k in lb if k.endswith('membrane.V')]
    assert maxdiff(la, lb, vk) < 1e-9


def test_ref_no_diffusion_field():
    m, _, _ = myokit.load('example')
    p = myokit.pacing.blocktrain(**PULSE)
    g = np.linspace(4, 16, 9)

    def setup(o):
        o.set_field('ina.gNa', g)
    w, ws, la = pair(m, p, 9, 6, ['engine.time', 'membrane.V', 'ina.INa'],
                     setup, diffusion=False)
    assert w == 0 and ws == 0
    # every cell is paced and they differ through the field only
    assert la['0.membrane.V'].max() > 0
    assert not np.array_equal(la['0.membrane.V'], la['8.membrane.V'])


# ---------------------------------------------------------------------------
# 3. Pacing
# ---------------------------------------------------------------------------
def test_pacing_known_answers():
    # myokit/tests/test_pacing_system_c.py:21-116 (values restated)
    # one event: level 1, start 100, duration 2, periodic 1000
    ev = [(1.0, 100.0, 2.0, 1000.0, 0.0)]
    times = [0, 50, 100, 101, 102, 500, 1100, 1101.9, 1102, 2100]
    levels, tnext = pacing_probe(ev, times)
    assert list(levels) == [0, 0, 1, 1, 0, 0, 1, 1, 0, 1]
    assert list(tnext) == [100, 100, 102, 102, 1100, 1100, 1102, 1102, 2100,
                           2102]
    # event at t = 0 fires immediately
    levels, tnext = pacing_probe([(2.0, 0.0, 1.0, 0.0, 0.0)], [0, 0.5, 1, 5])
    assert list(levels) == [2, 2, 0, 0]
    assert tnext[0] == 1 and np.isinf(tnext[2])
    # finite multiplier: exactly 3 occurrences
    levels, _ = pacing_probe([(1.0, 10.0, 1.0, 10.0, 3.0)],
                             [10, 20, 30, 40, 50])
    assert list(levels) == [1, 1, 1, 0, 0]
    # simultaneous events are an error
    with pytest.raises(RuntimeError):
        pacing_probe([(1.0, 10.0, 1.0, 0.0, 0.0), (2.0, 10.0, 1.0, 0.0, 0.0)],
                     [20])
    # negative start time
    levels, _ = pacing_probe([(1.0, -5.0, 10.0, 0.0, 0.0)], [-10, -5, 0, 5],
                             t0=-10)
    assert list(levels) == [0, 1, 1, 0]


def test_pacing_vs_python_pacing_system():
    rng = np.random.default_rng(11)
    for trial in range(20):
        p = myokit.Protocol()
        t = 0.0
        for k in range(rng.integers(1, 5)):
            t += float(rng.integers(1, 40)) * 0.5
            dur = float(rng.integers(1, 6)) * 0.25
            if k == 0 and rng.random() < 0.5:
                period = t + dur + 100.0
                p.schedule(float(rng.integers(1, 4)), t, dur, period, 0)
                break
            p.schedule(float(rng.integers(1, 4)), t, dur)
            t += dur
        ps = myokit.PacingSystem(p)
        ev = [(e.level(), e.start(), e.duration(), e.period(), e.multiplier())
              for e in p.events()]
        times = np.cumsum(rng.uniform(0, 3, size=200))
        want_l, want_n = [], []
        for tt in times:
            ps.advance(tt)
            want_l.append(ps.pace())
            want_n.append(ps.next_time())
        levels, tnext = pacing_probe(ev, times)
        assert list(levels) == want_l
        assert list(tnext) == want_n


# ---------------------------------------------------------------------------
# 4. Schedule
# ---------------------------------------------------------------------------
def test_schedule_counts_and_log_grid():
    m, p, _ = myokit.load('example')
    o = OracleSimulation(m, p, ncells=2, precision=DP)
    lg, _ = o.run(1000 * 0.005 * 40, log=['engine.time'], log_interval=1)
    # dt = 0.005, log_interval = 1: no intermediary steps
    assert o.last_steps == 40000
    assert np.allclose(lg['engine.time'], np.arange(0, 200, 1.0), atol=1e-9)
    # myokit/tests/test_simulation_log_interval.py:21-62
    o = OracleSimulation(m, p, ncells=2, precision=DP)
    lg, _ = o.run(10, log=['engine.time'], log_interval=0.5)
    assert len(lg['engine.time']) == 20
    assert np.max(np.abs(lg['engine.time'] - np.arange(0, 10, 0.5))) < 1e-2
    lg2, _ = o.run(10, log=['engine.time'], log_interval=0.5)
    assert np.max(np.abs(lg2['engine.time'] - np.arange(10, 20, 0.5))) < 1e-2
    # intermediary (sub-ulp) steps appear when k * dt and j * log_interval
    # disagree in floating point (SURVEY.md §7 hard parts): 0.1 vs 0.005
    o = OracleSimulation(m, p, ncells=2, precision=DP)
    lg, _ = o.run(10, log=['engine.time'], log_interval=0.1)
    assert len(lg['engine.time']) == 100
    assert o.last_steps > 2000
