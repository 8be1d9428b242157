"""
Pins the fibre-tissue oracle (oracle/fiber_tissue.py): this repository's
restatement against the reference's own rendered kernels (including its
``diff_step_fiber_tissue``, myokit/_sim/openclsim.cl:601-628) bit for bit, and
against the separately pinned single-grid oracle when the junction is open.
Models and sizes follow the reference's tests/test_simulation_fiber_tissue.py.
"""
import os

import numpy as np
import pytest

import myokit_b200     # noqa: F401  (locates myokit)
import myokit
from oracle.fiber_tissue import OracleFiberTissue
from oracle.oracle import OracleSimulation

DATA = os.path.join(os.path.dirname(myokit.__file__), 'tests', 'data')
DP = myokit.DOUBLE_PRECISION
SP = myokit.SINGLE_PRECISION


def models():
    mf = myokit.load_model(os.path.join(DATA, 'dn-1985-normalised.mmt'))
    mt = myokit.load_model(os.path.join(DATA, 'lr-1991.mmt'))
    return mf, mt


def make(kernel, precision=DP, gft=9.0, mf=None, mt=None):
    a, b = models()
    p = myokit.pacing.blocktrain(1000, 2.0, offset=.01)
    return OracleFiberTissue(
        mf or a, mt or b, p, ncells_fiber=(8, 4), ncells_tissue=(8, 6),
        nx_paced=4, g_fiber=(235, 100), g_tissue=(9, 5), g_fiber_tissue=gft,
        dt=0.0012, precision=precision, kernel=kernel,
        inter_log_tissue=['ina.INa'])


@pytest.mark.parametrize('precision', [DP, SP])
def test_port_equals_reference_rendered_kernels(precision):
    logf = ['membrane.V', 'membrane.i_diff']
    logt = ['membrane.V', 'membrane.i_diff', 'ina.INa']
    a = make('port', precision)
    ta, fa, tta = a.run(3.0, logf, logt, 0.25)
    b = make('ref', precision)
    tb, fb, ttb = b.run(3.0, logf, logt, 0.25)
    assert a.last_steps == b.last_steps and len(ta) == 12
    assert np.array_equal(ta, tb)
    for k in logf:
        assert np.array_equal(fa[k], fb[k])
    for k in logt:
        assert np.array_equal(tta[k], ttb[k])
    assert np.array_equal(a.fiber_state(), b.fiber_state())
    assert np.array_equal(a.tissue_state(), b.tissue_state())
    # the paced fibre fires and drives the tissue through the junction
    assert fa['membrane.V'].max() > 0 and tta['membrane.V'].max() > 0
    # junction rows of the tissue (cty = 1: rows 1..4) lead the others
    v = tta['membrane.V']
    first = [int(np.argmax(v[:, y, 0] > -40)) if (v[:, y, 0] > -40).any() else 99
             for y in range(6)]
    assert min(first[1:5]) <= min(first[0], first[5])


def test_open_junction_gives_two_independent_grids():
    mf, mt = models()
    p = myokit.pacing.blocktrain(1000, 2.0, offset=.01)
    ft = make('port', gft=0.0)
    t, f, tt = ft.run(2.0, ['membrane.V'], ['membrane.V'], 0.25)
    # fibre alone: the pinned single-grid oracle with the same paced rectangle
    o = OracleSimulation(mf, p, ncells=(8, 4), precision=DP)
    o.set_conductance(235, 100)
    o.set_paced_cells(4, 4, 0, 0)
    o.set_step_size(0.0012)
    log, state = o.run(2.0, log=['engine.time', 'membrane.V'], log_interval=0.25)
    assert np.array_equal(t, np.asarray(log['engine.time']))
    want = np.array([[[log['%d.%d.membrane.V' % (x, y)][k] for x in range(8)]
                      for y in range(4)] for k in range(len(t))])
    assert np.array_equal(f['membrane.V'], want)
    assert np.array_equal(ft.fiber_state(), np.asarray(state))
    # tissue alone and unpaced
    o = OracleSimulation(mt, p, ncells=(8, 6), precision=DP)
    o.set_conductance(9, 5)
    o.set_paced_cells(0, 0, 0, 0)
    o.set_step_size(0.0012)
    log, state = o.run(2.0, log=['engine.time', 'membrane.V'], log_interval=0.25)
    assert np.array_equal(ft.tissue_state(), np.asarray(state))


def test_same_model_both_sides_and_argument_checks():
    mf, mt = models()
    a = make('port', mf=mt, mt=mt)
    t, f, tt = a.run(1.0, ['membrane.V'], ['membrane.V'], 0.5)
    assert f['membrane.V'].shape == (2, 4, 8) and tt['membrane.V'].shape == (2, 6, 8)
    with pytest.raises(ValueError):
        OracleFiberTissue(mt, mt, None, ncells_fiber=(4, 8), ncells_tissue=(4, 6))
