"""
find_nan / NaN reporting on the device path, after the reference's own suite
(myokit/tests/test_simulation_opencl.py:1447-1589): same models, same expected
times, cells, variables and messages.
"""
import numpy as np
import pytest

import myokit_b200
import myokit

pytestmark = pytest.mark.gpu


def clamped_model():
    # LR1991 with V driven by the pacing signal ("voltage clamp"); the
    # calcium concentration takes the membrane_potential label so that the
    # grid still has something to diffuse. 1 / (1 - exp(0)) at V = -47.13.
    m = myokit.load_model('example')
    m.binding('pace').set_binding(None)
    v = m.get('membrane.V')
    v.set_rhs(-80)
    v.demote()
    v.set_label(None)
    v.set_binding('pace')
    m.get('ica.Ca_i').set_label('membrane_potential')
    m.get('membrane').move_variable(v, m.get('engine'))
    return m


def step_protocol(t, level0=-80):
    p = myokit.Protocol()
    p.schedule(start=0, level=level0, duration=t)
    p.schedule(start=t, level=-47.13, duration=1000)
    return p


def clamped_sim(protocol):
    s = myokit_b200.SimulationCUDA(clamped_model(), protocol, ncells=(3, 3))
    s.set_paced_cell_list([[1, 1]])
    return s


def test_huge_step_raises_and_needs_full_log():
    m = myokit.load_model('example')
    p = myokit.pacing.blocktrain(period=1000, duration=2)
    s = myokit_b200.SimulationCUDA(m, p, ncells=10)
    s.set_step_size(1)
    with pytest.raises(myokit.SimulationError):
        s.run(10)
    for log in (['membrane.V'], myokit.LOG_STATE, myokit.LOG_BOUND):
        s.reset()
        with pytest.raises(myokit.SimulationError, match='Unable to pinpoint'):
            s.run(10, log=log)


def test_one_over_zero_is_located_in_time():
    s = clamped_sim(step_protocol(1.234))
    with pytest.raises(myokit.SimulationError, match='Time:  1.23') as e:
        s.run(5)
    text = str(e.value)
    assert 'in cell (1,1)' in text
    assert 'IS paced' in text
    assert 'State before:' in text and 'Connected cells: (0,1), (2,1), (1,0), (1,2)' in text


def test_error_in_first_logged_point():
    s = clamped_sim(step_protocol(1.234))
    x = s.state(0, 0)
    x[3] = float('nan')
    s.set_state(x, 2, 2)
    with pytest.raises(myokit.SimulationError, match='met in the very first'):
        s.run(2)


def test_log_without_error():
    s = clamped_sim(None)
    d = s.run(10)
    with pytest.raises(myokit.FindNanError, match='not found in log'):
        s.find_nan(d)


def test_manual_call_and_watch_variable():
    s = clamped_sim(step_protocol(4, level0=-90))
    x = s.state(0, 0)
    x[2] = 0.9          # lower j in one cell, for the watch-variable search
    s.set_state(x, 2, 2)
    before = s.state()
    d = s.run(10, report_nan=False)
    after_time, after_state = s.time(), s.state()
    time, icell, variable, value, states, bounds = s.find_nan(d)
    assert abs(time - 4.005) < 1e-5
    assert icell == [1, 1]
    assert variable == 'ina.m'
    # `value` is the variable's last finite value, states[0] the bad point
    assert np.isfinite(value) and not np.all(np.isfinite(states[0]))
    assert 1 <= len(states) <= 4 and len(bounds) == len(states)
    # the search must leave the simulation as it found it
    assert s.time() == after_time
    assert np.array_equal(np.array(s.state()), np.array(after_state),
                          equal_nan=True)
    assert before != after_state

    with pytest.raises(myokit.FindNanError, match='not found'):
        s.find_nan(d, 'x.y', [0, 1])
    with pytest.raises(myokit.FindNanError, match='state'):
        s.find_nan(d, 'engine.time', [0, 1])
    with pytest.raises(myokit.FindNanError, match='safe range'):
        s.find_nan(d, 'ina.m')
    with pytest.raises(myokit.FindNanError, match='lower than'):
        s.find_nan(d, 'ina.m', [1, 0])

    time, icell, variable, value, states, bounds = s.find_nan(
        d, 'ina.j', [0.5, 1])
    assert 2 < time < 3
    assert icell == [2, 2] and variable == 'ina.j'
    time, icell, variable, value, states, bounds = s.find_nan(
        d, 'ina.j', [0.2, 1])
    assert 5 < time < 6
    assert icell == [2, 2] and variable == 'ina.j'

    # the 7-tuple form carries a log of every variable of every cell
    out = s.find_nan(d, return_log=True)
    assert len(out) == 7
    assert '1.1.ina.INa' in out[6] and '0.2.membrane.i_ion' in out[6]
