"""
The persistent kernel on the device (opt-in ``set_kernel_options(persistent=
True)``): it equals the oracle bit for bit on the host
(tests/test_generated_kernel_host.py); the device side — run grouping in
sim_step_typed — first ran (and passed) on a B200 in round 2.
"""
import numpy as np
import pytest

import myokit_b200
import myokit

pytestmark = pytest.mark.gpu
DP = myokit.DOUBLE_PRECISION


def test_persistent_cable_equals_default_kernel_and_oracle():
    m, p, _ = myokit.load('example')

    def make(persistent):
        s = myokit_b200.SimulationCUDA(m, p, ncells=128, precision=DP)
        s.set_conductance(10)
        s.set_paced_cells(5)
        s.set_kernel_options(persistent=persistent, fmad=False)
        return s
    a, b = make(True), make(False)
    log = ['engine.time', 'membrane.V', 'membrane.i_diff', 'ina.INa']
    da = a.run(80, log=log, log_interval=1)
    db = b.run(80, log=log, log_interval=1)
    assert np.max(np.asarray(da['0.membrane.V'])) > 0
    for k in db.keys():
        assert np.array_equal(np.asarray(da[k]), np.asarray(db[k])), k
    assert np.array_equal(a.state_array(), b.state_array())
    ia, ib = a.last_run_info(), b.last_run_info()
    assert ia['steps'] == ib['steps'] == 16000
    # (per logged row: one single-step launch, its two gather launches and one
    # launch for the 199 unlogged steps up to the next row)
    assert ia['kernel_launches'] < ib['kernel_launches'] // 40
    # a second run continues on the resident state
    da = a.run(20, log=['membrane.V'], log_interval=1)
    db = b.run(20, log=['membrane.V'], log_interval=1)
    assert np.array_equal(np.asarray(da['64.membrane.V']), np.asarray(db['64.membrane.V']))


def test_persistent_small_grid_rush_larsen_fields():
    m, _, _ = myokit.load('example')
    p = myokit.pacing.blocktrain(period=1000, duration=2, offset=1)
    rng = np.random.default_rng(2)
    gxf = rng.uniform(2, 9, size=(6, 9))
    gyf = rng.uniform(2, 9, size=(5, 10))

    def make(persistent):
        s = myokit_b200.SimulationCUDA(m, p, ncells=(10, 6), precision=DP, rl=True)
        s.set_conductance_field(gxf, gyf)
        s.set_paced_cells(3, 6, 0, 0)
        s.set_kernel_options(persistent=persistent, fmad=False)
        return s
    a, b = make(True), make(False)
    ta, fa = a.run_fields(6, ['membrane.V', 'membrane.i_diff'], log_interval=0.5)
    tb, fb = b.run_fields(6, ['membrane.V', 'membrane.i_diff'], log_interval=0.5)
    assert np.array_equal(ta, tb)
    for k in fb:
        assert np.array_equal(fa[k], fb[k])
    assert np.array_equal(a.state_array(), b.state_array())


def test_split_gates_equals_single_kernel_on_device():
    # kernelgen split_gates: mkb_cell_step + mkb_gate_step per step (also inside
    # the 64-step graphs) against the single kernel; same expressions, so the
    # results must be identical
    from myokit_b200 import workloads

    def make(split):
        s = workloads.c3_hetero(myokit_b200.SimulationCUDA, nx=96, ny=40)
        # (no FMA contraction: the compiler contracts per kernel, and the two
        # forms are different kernels)
        s.set_kernel_options(split_gates=split, fmad=False)
        return s
    a, b = make(True), make(False)
    assert a.kernel_source().gate_kernel
    ta, fa = a.run_fields(3.0, ['membrane.V', 'ina.m'], log_interval=0.5)
    tb, fb = b.run_fields(3.0, ['membrane.V', 'ina.m'], log_interval=0.5)
    assert fb['membrane.V'].max() > 0
    for k in fb:
        assert np.array_equal(fa[k], fb[k]), k
    assert np.array_equal(a.state_array(), b.state_array())
    assert a.last_run_info()['kernel_launches'] >= 2 * a.last_run_info()['steps']
