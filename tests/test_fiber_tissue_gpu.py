"""
FiberTissueSimulationCUDA on the device against the fibre-tissue oracle.

The junction kernels already equal the oracle bit for bit on the host
(tests/test_generated_kernel_host.py); what these tests add is the device
path: mkb_sim_junction_connect / mkb_sim_step_pair (first run, and green, on
a B200 in round 2).
"""
import os

import numpy as np
import pytest

import myokit_b200
import myokit
from oracle.fiber_tissue import OracleFiberTissue

pytestmark = pytest.mark.gpu

DATA = os.path.join(os.path.dirname(myokit.__file__), 'tests', 'data')
DP = myokit.DOUBLE_PRECISION
ARGS = dict(ncells_fiber=(8, 4), ncells_tissue=(8, 6), nx_paced=4,
            g_fiber=(235, 100), g_tissue=(9, 5), g_fiber_tissue=9, dt=0.0012)


def models():
    mf = myokit.load_model(os.path.join(DATA, 'dn-1985-normalised.mmt'))
    mt = myokit.load_model(os.path.join(DATA, 'lr-1991.mmt'))
    return mf, mt


def field(log, name, nx, ny):
    nt = len(log['engine.time'])
    return np.array([[[log['%d.%d.%s' % (x, y, name)][k] for x in range(nx)]
                      for y in range(ny)] for k in range(nt)])


def test_pair_equals_oracle_fp64_and_continues():
    mf, mt = models()
    p = myokit.pacing.blocktrain(1000, 2.0, offset=.01)
    s = myokit_b200.FiberTissueSimulationCUDA(mf, mt, p, precision=DP, **ARGS)
    o = OracleFiberTissue(mf, mt, p, precision=DP, **ARGS)
    names = ['membrane.V', 'membrane.i_diff']
    for part in range(2):       # second run: re-armed on the resident states
        logf, logt = s.run(2.0, logf=['engine.time'] + names,
                           logt=['engine.time'] + names, log_interval=0.25)
        tt, of, ot = o.run(2.0, names, names, 0.25)
        assert np.array_equal(np.asarray(logf['engine.time']), tt)
        assert np.array_equal(np.asarray(logt['engine.time']), tt)
        for name in names:
            assert np.abs(field(logf, name, 8, 4) - of[name]).max() <= 1e-6
            assert np.abs(field(logt, name, 8, 6) - ot[name]).max() <= 1e-6
    assert ot['membrane.V'].max() > 0           # the tissue was driven
    assert s.time() == 4.0
    assert np.abs(np.asarray(s.fiber_state()) - o.fiber_state()).max() <= 1e-6
    assert np.abs(np.asarray(s.tissue_state()) - o.tissue_state()).max() <= 1e-6


def test_pair_pre_reset_and_fp32():
    mf, mt = models()
    p = myokit.pacing.blocktrain(1000, 2.0, offset=.01)
    s = myokit_b200.FiberTissueSimulationCUDA(
        mf, mt, p, precision=myokit.SINGLE_PRECISION, **ARGS)
    s.pre(1.0)
    assert s.time() == 0
    d0 = np.asarray(s.default_fiber_state())
    assert np.array_equal(d0, np.asarray(s.fiber_state()))
    logf, logt = s.run(1.0, logf=['membrane.V'], logt=['membrane.V'],
                       log_interval=0.5)
    assert len(logf['0.0.membrane.V']) == 2
    s.reset()
    assert s.time() == 0 and np.array_equal(np.asarray(s.fiber_state()), d0)
    o = OracleFiberTissue(mf, mt, p, precision=myokit.SINGLE_PRECISION, **ARGS)
    o.run(1.0, [], [], 1.0)
    o.set_time(0)           # pre() leaves the time where it was
    tt, of, ot = o.run(1.0, ['membrane.V'], ['membrane.V'], 0.5)
    # fp32: close, not identical (fast division, FMA contraction)
    assert np.abs(field(dict(logf, **{'engine.time': tt}), 'membrane.V', 8, 4)
                  - of['membrane.V']).max() <= 0.5
