"""
Parity at the sizes SURVEY.md §8(d) prescribes (the smaller cases in
test_parity_gpu.py run first): C2 512 x 512 fp32 with row 256 logged every
step for 60 ms, C3 as a 256 x 256 crop with the bench's seeds, C4-proxy
512 x 512 fp32 for 20 ms, and the double-precision bar "max |dV| <= 1e-6 mV
per logged sample over 1 s" on a 2-d grid. The oracle runs with OpenMP over
the host cores (cells are independent within a pass, so the thread count does
not change its results).
"""
import os

import numpy as np
import pytest

import myokit_b200
from myokit_b200 import workloads
import myokit

from oracle.oracle import OracleSimulation
from util import max_abs_diff

pytestmark = pytest.mark.gpu
DP = myokit.DOUBLE_PRECISION
SP = myokit.SINGLE_PRECISION
CORES = os.cpu_count() or 1


def as_arrays(log):
    log = log[0] if isinstance(log, tuple) else log
    return dict((k, np.asarray(v, dtype=np.float64)) for k, v in log.items())


def activation_times(log, keys, times, threshold=-20.0):
    out = []
    for k in keys:
        idx = np.nonzero(np.asarray(log[k]) > threshold)[0]
        out.append(times[idx[0]] if len(idx) else np.nan)
    return np.array(out)


def test_c2_512_row_256_activation_times_over_60_ms():
    n, dt = 512, 0.005
    keys = ['%d.256.membrane.V' % x for x in range(n)]
    logspec = ['engine.time'] + keys
    a = workloads.c2_planar(myokit_b200.SimulationCUDA, n)
    la = as_arrays(a.run(60, log=logspec, log_interval=dt))
    b = workloads.c2_planar(OracleSimulation, n, openmp=True)
    lb = as_arrays(b.run(60, log=logspec, log_interval=dt, nthreads=CORES))
    assert np.array_equal(la['engine.time'], lb['engine.time'])
    assert len(la['engine.time']) == 12000
    ta = activation_times(la, keys, la['engine.time'])
    tb = activation_times(lb, keys, lb['engine.time'])
    crossed = ~np.isnan(tb)
    assert crossed.sum() >= 200, 'the wave must have travelled'
    assert np.array_equal(np.isnan(ta), np.isnan(tb))
    # north star: single precision, activation times within one dt
    assert np.max(np.abs(ta[crossed] - tb[crossed])) <= dt * 1.0001
    # and the fp32 division in use (__fdividef, 2 ulp) keeps V itself close
    assert max_abs_diff(la, lb, keys) < 0.5


def test_c3_256_crop_same_seeds():
    n = 256
    cells = ['3.%d.' % (n // 2), '100.17.', '128.128.', '255.255.', '0.0.']
    logspec = ['engine.time'] + [c + v for c in cells
                                 for v in ('membrane.V', 'membrane.i_diff')]
    a = workloads.c3_hetero(myokit_b200.SimulationCUDA, nx=n)
    la = as_arrays(a.run(8, log=logspec, log_interval=0.25))
    b = workloads.c3_hetero(OracleSimulation, nx=n, openmp=True)
    lb, sb = b.run(8, log=logspec, log_interval=0.25, nthreads=CORES)
    lb = as_arrays(lb)
    assert la['3.%d.membrane.V' % (n // 2)].max() > 0      # paced edge fired
    assert max_abs_diff(la, lb, ['engine.time']) == 0
    assert max_abs_diff(la, lb, suffix='membrane.V') <= 1e-6    # north star, fp64
    assert max_abs_diff(la, lb, suffix='i_diff') <= 1e-6
    sa = a.state_array()
    scale = np.abs(sb).reshape(-1, a._nstate).max(axis=0) + 1e-30
    rel = np.abs(sa - sb).reshape(-1, a._nstate) / scale
    assert rel.max() <= 1e-9, rel.max()


def test_c4_proxy_512_fp32_20_ms_state_within_1e3():
    n = 512
    m, _, _ = myokit.load('example')
    ns = m.count_states()
    init = np.array(m.initial_values(True))
    state = np.tile(init, n * n).reshape(n, n, ns)
    iv = m.get('membrane.V').index()
    ih, ij = m.get('ina.h').index(), m.get('ina.j').index()
    # broken wave: depolarised band in the lower half, refractory tail beside it
    # (the tail starts below -47.13 mV: the model's ina.m.alpha is 0 / 0 there,
    # and cells that relax slowly through that value in single precision hit it
    # — in this back-end and in the oracle alike: NaN in 4976 cells after 5 ms)
    state[:n // 2, 40:60, iv] = 10.0
    state[:n // 2, 0:40, ih] = 0.0
    state[:n // 2, 0:40, ij] = 0.0
    state[:n // 2, 0:40, iv] = -55.0
    res = []
    for cls, kw, run_kw in ((myokit_b200.SimulationCUDA, {}, {}),
                            (OracleSimulation, dict(openmp=True), dict(nthreads=CORES))):
        s = cls(m, None, ncells=(n, n), precision=SP, **kw)
        s.set_conductance(1, 1)
        s.set_paced_cells(0, 0, 0, 0)
        s.set_step_size(0.005)
        s.set_state(state.ravel())
        r = s.run(20, log=['engine.time', '300.100.membrane.V'], log_interval=1,
                  **run_kw)
        res.append(np.asarray(r[1]) if isinstance(r, tuple) else s.state_array())
    sa, sb = res
    V = sa.reshape(n, n, ns)[:, :, iv]
    assert V.max() - V.min() > 50           # a wave is there
    scale = np.abs(sb).reshape(-1, ns).max(axis=0)
    rel = np.abs(sa - sb).reshape(-1, ns) / scale
    assert rel.max() <= 1e-3, rel.max()


def test_fp64_two_dimensional_one_second():
    # 64 x 64 LR1991, 1 Hz pacing of a corner patch, V of every 8th cell logged
    # every ms for 1000 ms: max |dV| <= 1e-6 mV per logged sample
    n = 64
    m, p, _ = myokit.load('example')
    keys = ['%d.%d.membrane.V' % (x, y) for y in range(0, n, 8) for x in range(0, n, 8)]
    logspec = ['engine.time'] + keys
    res = []
    for cls, kw, run_kw in ((myokit_b200.SimulationCUDA, {}, {}),
                            (OracleSimulation, dict(openmp=True), dict(nthreads=CORES))):
        s = cls(m, p, ncells=(n, n), precision=DP, **kw)
        s.set_conductance(10, 5)
        s.set_paced_cells(6, 6, 0, 0)
        s.set_step_size(0.005)
        res.append(as_arrays(s.run(1000, log=logspec, log_interval=1, **run_kw)))
    la, lb = res
    assert len(la['engine.time']) == 1000
    assert np.array_equal(la['engine.time'], lb['engine.time'])
    assert la['56.56.membrane.V'].max() > 0         # the far corner fired
    w = max_abs_diff(la, lb, keys)
    assert w <= 1e-6, w
