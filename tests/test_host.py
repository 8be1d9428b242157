"""
Host-side tests that need no GPU: the C-ABI library loads and exports what
include/myokit_b200.h declares, the JIT cross-compiles every kernel variant
for sm_100a, the pacing system matches Myokit's, SimulationCUDA validates its
arguments like the reference class (myokit/tests/test_simulation_opencl.py),
and — on a machine without a GPU — running fails loudly instead of falling
back to anything.
"""
import ctypes
import os
import re

import numpy as np
import pytest

import myokit_b200
from myokit_b200 import capi, workloads
import myokit

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DP = myokit.DOUBLE_PRECISION
SP = myokit.SINGLE_PRECISION
HAS_GPU = capi.device_count() > 0


def br():
    return workloads.data_model('beeler-1977-model.mmt')


# ---------------------------------------------------------------------------
# C ABI
# ---------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    with open(os.path.join(ROOT, 'include', 'myokit_b200.h')) as f:
        text = f.read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    declared = set(re.findall(r'\b(mkb_[a-z_0-9]+)\s*\(', text))
    assert len(declared) >= 15
    lib = capi.library()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(capi.SYMBOLS)
    assert lib.mkb_abi_version() == capi.MKB_ABI_VERSION


def test_config_struct_matches_header_size():
    # Guard against the ctypes mirror drifting from the C struct: compile a
    # one-liner with the real header and compare sizeof.
    import subprocess
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, 's.c')
        with open(src, 'w') as f:
            f.write('#include <stdio.h>\n#include "myokit_b200.h"\n'
                    'int main(){printf("%zu %zu %zu", sizeof(mkb_sim_config), '
                    'sizeof(mkb_device_info_t), sizeof(mkb_run_config));'
                    'return 0;}')
        exe = os.path.join(d, 's')
        subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'),
                               src, '-o', exe])
        out = subprocess.check_output([exe]).decode().split()
    assert int(out[0]) == ctypes.sizeof(capi.SimConfig)
    assert int(out[1]) == ctypes.sizeof(capi.DeviceInfo)
    assert int(out[2]) == ctypes.sizeof(capi.RunConfig)


def test_device_abi_header_is_embedded():
    text = capi.library().mkb_device_abi_header().decode()
    assert 'struct MkbGridArgs' in text and 'struct MkbStepParams' in text
    with open(os.path.join(ROOT, 'myokit_b200', 'csrc',
                           'mkb_device_abi.h')) as f:
        assert f.read() == text


def test_jit_reports_errors():
    with pytest.raises(capi.BackendError) as e:
        capi.jit_compile('this is not CUDA')
    assert e.value.code == capi.MKB_ERR_JIT


def test_pacing_probe_matches_myokit_pacing_system():
    lib = capi.library()
    rng = np.random.default_rng(3)
    for trial in range(20):
        p = myokit.Protocol()
        t = 0.0
        for k in range(rng.integers(1, 5)):
            t += float(rng.integers(1, 40)) * 0.5
            dur = float(rng.integers(1, 6)) * 0.25
            if k == 0 and rng.random() < 0.5:
                p.schedule(1.5, t, dur, t + dur + 50.0, int(rng.integers(0, 4)))
                break
            p.schedule(float(rng.integers(1, 4)), t, dur)
            t += dur
        ps = myokit.PacingSystem(p)
        ev = np.array([[e.level(), e.start(), e.duration(), e.period(),
                        e.multiplier()] for e in p.events()]).ravel()
        times = np.cumsum(rng.uniform(0, 3, size=300))
        want_l, want_n = [], []
        for tt in times:
            ps.advance(tt)
            want_l.append(ps.pace())
            want_n.append(ps.next_time())
        levels = np.zeros(len(times))
        tnext = np.zeros(len(times))
        rc = lib.mkb_pacing_probe(
            0.0, len(ev) // 5, ev.ctypes.data, len(times), times.ctypes.data,
            levels.ctypes.data, tnext.ctypes.data)
        assert rc == 0
        assert list(levels) == want_l
        assert list(tnext) == want_n


def schedule(tmin, tmax, dt, log_interval, events, max_steps=10 ** 6):
    lib = capi.library()
    ev = np.array(events, dtype=np.float64).ravel()
    n_ev = len(ev) // 5
    if n_ev == 0:
        ev = np.zeros(5)
    times = np.zeros(max_steps)
    dts = np.zeros(max_steps)
    paces = np.zeros(max_steps)
    logging = np.zeros(max_steps, dtype=np.uint8)
    n = ctypes.c_uint64(0)
    rc = lib.mkb_schedule_probe(
        tmin, tmax, dt, log_interval, n_ev, ev.ctypes.data, max_steps,
        times.ctypes.data, dts.ctypes.data, paces.ctypes.data,
        logging.ctypes.data, ctypes.byref(n))
    assert rc == 0, lib.mkb_last_error()
    n = n.value
    return times[:n], dts[:n], paces[:n], logging[:n].astype(bool)


def test_schedule_matches_the_oracle_loop():
    # The runtime's step selection (mkb_schedule.hpp) against the oracle's
    # independent restatement of openclsim.c:1051-1178, on the cases where
    # the loop is subtle: sub-ulp intermediary steps, pacing events off the
    # step grid, negative start times, log every step.
    from oracle.oracle import OracleSimulation
    m, _, _ = myokit.load('example')
    cases = [
        (0.0, 40.0, 0.005, 1.0, [(1, 10.0, 0.5, 0, 0)]),
        (0.0, 30.0, 0.005, 0.1, [(1, 3.0, 2.0, 7.0, 0)]),     # 0.1 vs 0.005: micro-steps
        (0.0, 20.0, 0.01, 0.3, [(2, 1.234, 0.777, 0, 0)]),     # event off the grid
        (-3.0, 9.0, 0.005, 0.5, [(1, 0.0, 1.0, 4.0, 2)]),      # negative start, 2 repeats
        (5.0, 6.5, 0.02, 1e-9, []),                             # log every step
        (0.0, 10.0, 0.005, 2.5, [(1, 0.0, 0.5, 1.0, 0)]),      # event at t = 0
    ]
    for tmin, tmax, dt, li, events in cases:
        t, dts, pace, logging = schedule(tmin, tmax, dt, li, events)
        assert abs(t[-1] + dts[-1] - tmax) < 1e-9 or t[-1] + dts[-1] >= tmax
        p = myokit.Protocol()
        for level, start, dur, period, mult in events:
            p.schedule(level, start, dur, period, mult)
        o = OracleSimulation(m, p, ncells=1, precision=DP)
        o.set_step_size(dt)
        o.set_time(tmin)
        lg, _ = o.run(tmax - tmin, log=['engine.time', 'engine.pace'],
                      log_interval=li)
        assert len(t) == o.last_steps, (tmin, tmax, dt, li)
        assert np.array_equal(t[logging], lg['engine.time'])
        assert np.array_equal(pace[logging], lg['engine.pace'])
    # properties of the loop itself
    t, dts, pace, logging = schedule(0.0, 1000.0, 0.005, 1.0,
                                     [(1, 50.0, 0.5, 1000.0, 0)])
    assert len(t) == 200000 and logging.sum() == 1000
    assert np.all(dts > 0) and pace.max() == 1
    t, dts, pace, logging = schedule(0.0, 100.0, 0.005, 0.1, [])
    assert logging.sum() == 1000 and len(t) > 20000          # intermediary steps
    assert dts.min() < 1e-12


def test_pacing_probe_simultaneous_events():
    lib = capi.library()
    ev = np.array([1.0, 10, 1, 0, 0, 2.0, 10, 1, 0, 0])
    times = np.array([20.0])
    out = np.zeros(1)
    rc = lib.mkb_pacing_probe(0.0, 2, ev.ctypes.data, 1, times.ctypes.data,
                              out.ctypes.data, out.ctypes.data)
    assert rc == capi.MKB_ERR_SIMULTANEOUS
    assert b'same time' in lib.mkb_last_error()


# ---------------------------------------------------------------------------
# Kernel generator
# ---------------------------------------------------------------------------
def variants():
    m, p, _ = myokit.load('example')
    S = myokit_b200.SimulationCUDA
    out = {}
    out['1d_fp64'] = workloads.c1_cable(S, 64)
    out['2d_fp32'] = workloads.c2_planar(S, 32)
    out['2d_hetero_rl_field'] = workloads.c3_hetero(S, nx=16)
    s = S(m, p, ncells=16, precision=DP)
    s.set_connections([(0, 1, 1.0), (1, 5, 2.0)])
    out['connections'] = s
    s = S(m, p, ncells=8, diffusion=False, precision=SP, native_maths=True)
    out['no_diffusion_native'] = s
    s = S(m, p, ncells=(8, 8), precision=DP)
    s.set_paced_cell_list([(0, 0), (3, 4)])
    out['paced_list'] = s
    return out


def test_every_kernel_variant_compiles_for_sm100a():
    for name, s in variants().items():
        src = s.kernel_source()
        cubin, log = capi.jit_compile(src.code, src.options)
        assert cubin[:4] == b'\x7fELF', name
        assert src.kernel_name.encode() in cubin, name


def test_generated_source_follows_reference_arithmetic():
    s = variants()['1d_fp64']
    # tiny models default to several cells per thread with vector accesses;
    # any model can ask for it
    s.set_kernel_options(cells_per_thread=2, persistent=False)
    code = s.kernel_source().code
    assert '#define MKB_CPT 2' in code
    assert 'mkb_vload<MKB_CPT>(S1[r], state + 1ull * stride + cid0, active);' in code
    assert 'N1[c] = V_m + dt * D_m;' in code
    assert 'mkb_vstore<MKB_CPT>(v_out + cid0, N0);' in code
    assert 'gx * (2 * vcc - vxm - vxp)' in code
    # the one-cell-per-thread form
    s.set_kernel_options(cells_per_thread=1)
    code = s.kernel_source().code
    # expression text comes from myokit's own CUDA writer; integer powers
    # become multiplication chains unless asked otherwise
    assert 'mkb_powi<3>(V_m)' in code
    s.set_kernel_options(pow_multiply=False)
    assert 'pow(V_m, 3.0)' in s.kernel_source().code
    # zero-flux stencil forms, openclsim.cl:406-415
    assert 'gx * (vc - vxp)' in code and 'gx * (2 * vc - vxm - vxp)' in code
    # forward Euler update, V goes to the second V plane
    assert 'v_out[cid] = V_V + dt * D_V;' in code
    assert 'MKB_AT(state_c, 1) = V_m + dt * D_m;' in code
    # Rush-Larsen update, openclsim.cl:362
    code = variants()['2d_hetero_rl_field'].kernel_source().code
    # (tau = 1 / X: the exponent -dt / tau is written -dt * X)
    assert re.search(r'= V_\w+ - \(V_\w+ - V_\w+\) \* mkb_exp_poly\(\(-dt \* \(V_\w+ \+ V_\w+\)\)\);', code)
    s = variants()['2d_hetero_rl_field']
    s.set_kernel_options(const_div=False)
    assert re.search(r'= V_\w+ - \(V_\w+ - V_\w+\) \* mkb_exp_poly\(mkb_div\(-dt, V_\w+\)\);',
                     s.kernel_source().code)
    s = variants()['2d_hetero_rl_field']
    s.set_kernel_options(fast_div=False, fast_exp=False)
    assert re.search(r'= V_\w+ - \(V_\w+ - V_\w+\) \* exp\(-dt / V_\w+\);',
                     s.kernel_source().code)
    assert 'gxm = has_xm ? gxf[cid - iy - 1]' in code
    assert 'idiff += gxm * (vc - vxm)' in code
    # fp32: float literals and float maths
    code = variants()['2d_fp32'].kernel_source().code
    assert 'typedef float Real;' in code and 'expf(' in code
    assert re.search(r'\d\.\d+f\b', code)
    # native maths
    code = variants()['no_diffusion_native'].kernel_source().code
    assert '__expf(' in code


def test_logged_intermediaries_are_stored_only_on_logged_steps():
    m, p, _ = myokit.load('example')
    s = myokit_b200.SimulationCUDA(m, p, ncells=8, precision=DP)
    s.set_kernel_options(persistent=False)
    src = s.kernel_source([s._model.get('ica.ICa')])
    assert src.n_inter == 1
    assert 'if (store_aux) MKB_AT(inter_c, 0) = V_ICa;' \
        in src.code


# ---------------------------------------------------------------------------
# SimulationCUDA argument checking (reference: test_simulation_opencl.py)
# ---------------------------------------------------------------------------
def test_creation_errors():
    m = br()
    S = myokit_b200.SimulationCUDA
    m2 = m.clone()
    m2.label('membrane_potential').set_rhs(None)
    with pytest.raises(myokit.MissingRhsError):
        S(m2)
    m2 = m.clone()
    x = m2.get('ix1').add_variable('xx')
    x.set_rhs('membrane.i_ion')
    with pytest.raises(ValueError, match='interdependent'):
        S(m2)
    assert S(m, ncells=1).shape() == 1
    assert S(m, ncells=50).shape() == 50
    assert S(m, ncells=(2, 1)).shape() == (1, 2)
    for bad in (None, (1,), (1, 1, 1)):
        with pytest.raises(ValueError, match=r'scalar or a tuple \(nx, ny\)'):
            S(m, ncells=bad)
    for bad in (-1, 0, (0, 10), (10, 0), (-1, -1)):
        with pytest.raises(ValueError, match='at least 1'):
            S(m, ncells=bad)
    with pytest.raises(ValueError, match='Only single and double'):
        S(m, precision=SP + DP)
    m2 = m.clone()
    m2.label('membrane_potential').set_label(None)
    with pytest.raises(ValueError, match='requires the membrane potential'):
        S(m2)
    m2.get('ina.INa').set_label('membrane_potential')
    with pytest.raises(ValueError, match='must be a state variable'):
        S(m2)
    # without diffusion or RL the label is not needed
    m2 = m.clone()
    m2.label('membrane_potential').set_label(None)
    S(m2, diffusion=False)


def test_conductance_setters():
    m = br()
    s = myokit_b200.SimulationCUDA(m, ncells=4)
    assert s.conductance() == 10
    s.set_conductance(3)
    assert s.conductance() == 3
    with pytest.raises(ValueError, match='Invalid conductance gx'):
        s.set_conductance(-1)
    s2 = myokit_b200.SimulationCUDA(m, ncells=(4, 3))
    assert s2.conductance() == (10, 5)
    with pytest.raises(ValueError, match='Invalid conductance gy'):
        s2.set_conductance(1, -1)
    # fields: shapes (nx-1,) / (ny, nx-1), (ny-1, nx)
    s.set_conductance_field([1, 2, 3])
    assert s.conductance() is None
    with pytest.raises(ValueError, match='must have length 3'):
        s.set_conductance_field([1, 2])
    with pytest.raises(ValueError, match='must be None'):
        s.set_conductance_field([1, 2, 3], [1])
    with pytest.raises(ValueError, match='negative'):
        s.set_conductance_field([1, -2, 3])
    s2.set_conductance_field(np.ones((3, 3)), np.ones((2, 4)))
    assert s2.conductance() is None
    with pytest.raises(ValueError, match=r'`gx` must have dimensions'):
        s2.set_conductance_field(np.ones((3, 4)), np.ones((2, 4)))
    with pytest.raises(ValueError, match=r'`gy` must be set'):
        s2.set_conductance_field(np.ones((3, 3)))
    with pytest.raises(ValueError, match=r'`gy` must have dimensions'):
        s2.set_conductance_field(np.ones((3, 3)), np.ones((3, 4)))
    with pytest.raises(ValueError, match='negative'):
        s2.set_conductance_field(np.ones((3, 3)), -np.ones((2, 4)))
    # set_conductance clears fields
    s.set_conductance(7)
    assert s.conductance() == 7
    # alias from BASELINE.json's north_star
    s.set_heterogeneity([1, 2, 3])
    assert s.conductance() is None


def test_connections_validation():
    m = br()
    s = myokit_b200.SimulationCUDA(m, ncells=4)
    s.set_connections([(0, 1, 1.0), (2, 1, 0.5)])
    assert s.conductance() is None
    assert s.neighbors(1) == [0, 2]
    with pytest.raises(ValueError, match='cannot be None'):
        s.set_connections(None)
    with pytest.raises(ValueError, match='list of 3-tuples'):
        s.set_connections([(0, 1)])
    for bad in [(0, 0, 1), (-1, 0, 1), (0, 4, 1), (0, -1, 1)]:
        with pytest.raises(ValueError, match='Invalid connection'):
            s.set_connections([bad])
    with pytest.raises(ValueError, match='Duplicate connection'):
        s.set_connections([(0, 1, 1), (1, 0, 1)])
    with pytest.raises(ValueError, match='Invalid conductance'):
        s.set_connections([(0, 1, -1)])
    s2 = myokit_b200.SimulationCUDA(m, ncells=(4, 3))
    with pytest.raises(RuntimeError, match='1d mode'):
        s2.set_connections([(0, 1, 1)])


def test_connections_as_arrays():
    m = br()
    s = myokit_b200.SimulationCUDA(m, ncells=6)
    i = np.array([0, 2, 5])
    j = np.array([1, 1, 0])
    g = np.array([1.0, 0.5, 2.0])
    s.set_connections((i, j, g))
    assert s.neighbors(1) == [0, 2] and s.neighbors(0) == [1, 5]
    s.set_connections(np.stack([i, j, g], axis=1))
    assert s.neighbors(5) == [0]
    for bad_i, bad_j, bad_g, msg in (
            ([0], [0], [1.0], 'Invalid connection'),
            ([0], [6], [1.0], 'Invalid connection'),
            ([-1], [2], [1.0], 'Invalid connection'),
            ([0, 1], [1, 0], [1.0, 1.0], 'Duplicate connection'),
            ([0], [1], [-1.0], 'Invalid conductance')):
        with pytest.raises(ValueError, match=msg):
            s.set_connections((np.array(bad_i), np.array(bad_j),
                               np.array(bad_g)))
    n, (ei, ej, eg) = workloads.fibre_mesh(8, 4, 3, extra=0.2)
    lattice = 7 * 4 * 3 + 8 * 3 * 3 + 8 * 4 * 2
    assert n == 96 and lattice <= len(ei) <= lattice + 19
    assert np.all(ei < ej) and len(np.unique(ei * n + ej)) == len(ei)
    big = myokit_b200.SimulationCUDA(m, ncells=n)
    big.set_connections((ei, ej, eg))


def test_diffusion_disabled_methods_raise():
    m = br()
    s = myokit_b200.SimulationCUDA(m, ncells=4, diffusion=False)
    for call in (s.conductance, lambda: s.is_paced(0), lambda: s.neighbors(0),
                 s.set_conductance, lambda: s.set_conductance_field([1] * 3),
                 lambda: s.set_connections([(0, 1, 1)]), s.set_paced_cells,
                 lambda: s.set_paced_cell_list([0])):
        with pytest.raises(RuntimeError, match='unavailable when diffusion'):
            call()


def test_paced_cells_and_neighbors():
    m = br()
    s = myokit_b200.SimulationCUDA(m, ncells=10)
    assert [s.is_paced(i) for i in range(10)] == [True] * 5 + [False] * 5
    s.set_paced_cells(2, x=3)
    assert [i for i in range(10) if s.is_paced(i)] == [3, 4]
    s.set_paced_cells(-2, x=5)          # left of x
    assert [i for i in range(10) if s.is_paced(i)] == [3, 4]
    s.set_paced_cells(2, x=-3)          # counted from the right
    assert [i for i in range(10) if s.is_paced(i)] == [7, 8]
    s.set_paced_cell_list([1, 9, 1])
    assert [i for i in range(10) if s.is_paced(i)] == [1, 9]
    with pytest.raises(IndexError):
        s.set_paced_cell_list([10])
    with pytest.raises(IndexError):
        s.is_paced(10)
    with pytest.raises(ValueError):
        s.is_paced(0, 1)
    assert s.neighbors(0) == [1]
    assert s.neighbors(4) == [3, 5]
    assert s.neighbors(9) == [8]
    s2 = myokit_b200.SimulationCUDA(m, ncells=(4, 3))
    assert s2.is_paced(3, 2)            # default 5 x 5 rectangle covers it
    s2.set_paced_cells(1, 2, 1, 1)
    assert [(x, y) for y in range(3) for x in range(4)
            if s2.is_paced(x, y)] == [(1, 1), (1, 2)]
    s2.set_paced_cell_list([(0, 0), (3, 2)])
    assert s2.is_paced(3, 2) and not s2.is_paced(2, 2)
    with pytest.raises(IndexError):
        s2.set_paced_cell_list([(4, 0)])
    with pytest.raises(ValueError):
        s2.is_paced(0)
    assert s2.neighbors(0, 0) == [(1, 0), (0, 1)]
    assert s2.neighbors(1, 1) == [(0, 1), (2, 1), (1, 0), (1, 2)]
    assert s2.neighbors(3, 2) == [(2, 2), (3, 1)]


def test_state_handling():
    m = br()
    n = m.count_states()
    s = myokit_b200.SimulationCUDA(m, ncells=(3, 2))
    init = m.initial_values(True)
    assert s.state() == init * 6
    assert s.state(2, 1) == init
    assert s.default_state(0, 0) == init
    one = [float(i) for i in range(n)]
    s.set_state(one, 1, 1)
    assert s.state(1, 1) == one
    assert s.state(0, 1) == init
    # x changes first: cell (1, 1) is cell 1 + 1 * 3 = 4
    assert s.state()[4 * n:5 * n] == one
    s.set_state(one)
    assert s.state() == one * 6
    full = list(np.arange(6 * n, dtype=float))
    s.set_state(full)
    assert s.state(2, 0) == full[2 * n:3 * n]
    with pytest.raises(ValueError, match='argument x'):
        s.set_state(full, 1)
    with pytest.raises(ValueError, match='same size'):
        s.set_state(one[:-1])
    with pytest.raises(IndexError):
        s.set_state(one, 3, 0)
    with pytest.raises(IndexError):
        s.state(0, 2)
    s.reset()
    assert s.state() == init * 6
    s.set_default_state(one, 0, 0)
    s.reset()
    assert s.state(0, 0) == one and s.state(1, 0) == init
    s1 = myokit_b200.SimulationCUDA(m, ncells=3)
    with pytest.raises(ValueError):
        s1.state(0, 1)


def test_fields_constants_time_step():
    m = br()
    s = myokit_b200.SimulationCUDA(m, ncells=(3, 2))
    s.set_field('ina.gNaBar', np.ones((2, 3)))
    with pytest.raises(ValueError, match='dimensions'):
        s.set_field('ina.gNaBar', np.ones((3, 2)))
    with pytest.raises(ValueError, match='Only constants'):
        s.set_field('membrane.V', np.ones((2, 3)))
    with pytest.raises(ValueError, match='Bound values'):
        s.set_field('engine.pace', np.ones((2, 3)))
    s.remove_field('ina.gNaBar')
    with pytest.raises(KeyError):
        s.remove_field('ina.gNaBar')
    s.set_constant('ina.gNaBar', 3.5)
    assert 'V_gNaBar = 3.5f' in s.kernel_source().code
    with pytest.raises(ValueError, match='not a literal'):
        s.set_constant('membrane.i_ion', 1)
    assert s.step_size() == 0.005
    s.set_step_size(0.01)
    assert s.step_size() == 0.01
    with pytest.raises(ValueError, match='greater than zero'):
        s.set_step_size(0)
    assert s.time() == 0
    s.set_time(3)
    assert s.time() == 3
    assert s.is_2d() and not myokit_b200.SimulationCUDA(m, ncells=3).is_2d()
    assert s.monodomain_conductance(2, 3, 5, 0.1) == pytest.approx(
        5 * 3 / (4 * 2 * 0.01))


def test_run_argument_checks_and_zero_duration():
    m = br()
    s = myokit_b200.SimulationCUDA(m, ncells=2)
    with pytest.raises(ValueError, match='negative'):
        s.run(-1)
    d = s.run(0)                # legal, never reaches the back-end
    assert len(d) == 2 * m.count_states() + 1 + 1 + 2
    with pytest.raises(ValueError, match='ProgressReporter'):
        s.run(1, progress=12)


@pytest.mark.skipif(HAS_GPU, reason='checks the no-GPU failure mode')
def test_no_gpu_means_loud_failure_not_fallback():
    m = br()
    s = myokit_b200.SimulationCUDA(m, ncells=2)
    with pytest.raises(capi.BackendError) as e:
        s.run(1)
    assert e.value.code == capi.MKB_ERR_CUDA
    assert 'no CPU fallback' in str(e.value)
    # and the product never touches the oracle
    import sys
    for name, mod in list(sys.modules.items()):
        if name.startswith('myokit_b200'):
            src = getattr(mod, '__file__', None)
            if src and src.endswith('.py'):
                with open(src) as f:
                    text = f.read()
                assert 'import oracle' not in text
                assert 'from oracle' not in text


def test_cuda_info_class(tmp_path, monkeypatch):
    from myokit_b200 import CUDA
    from myokit_b200 import cuda as cuda_mod
    assert CUDA.supported() == HAS_GPU
    assert isinstance(CUDA.info(formatted=True), str)
    assert len(CUDA.available()) == capi.device_count()
    monkeypatch.setattr(cuda_mod.CUDA, '_path',
                        staticmethod(lambda: str(tmp_path / 'sel.ini')))
    monkeypatch.delenv('MYOKIT_CUDA_DEVICE', raising=False)
    assert CUDA.load_selection() == 0
    CUDA.save_selection(3)
    assert CUDA.load_selection() == 3
    monkeypatch.setenv('MYOKIT_CUDA_DEVICE', '1')
    assert CUDA.load_selection() == 1
    monkeypatch.delenv('MYOKIT_CUDA_DEVICE')
    CUDA.save_selection(None)
    assert CUDA.load_selection() == 0
    assert cuda_mod.bytesize(2048) == '2 KB'
    assert cuda_mod.clockspeed(1.965e9) == '1.965 GHz'
    if not HAS_GPU:
        with pytest.raises(myokit_b200.NoCUDAError):
            CUDA.current_info()


def test_export_kernel_compiles_with_nvcc(tmp_path):
    import shutil
    import subprocess
    s = workloads.c3_hetero(myokit_b200.SimulationCUDA, nx=16)
    path = s.export_kernel(str(tmp_path))
    assert os.path.isfile(path)
    assert os.path.isfile(os.path.join(str(tmp_path), 'mkb_device_abi.h'))
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    with open(path) as f:
        first = f.readline()
    assert first.startswith('// Build: nvcc')
    cmd = first[len('// Build: '):].split()
    cmd[0] = nvcc
    # nvcc spells the option --fmad like NVRTC does; -default-device is implied
    r = subprocess.run(cmd, cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    assert os.path.getsize(os.path.join(str(tmp_path), 'mkb_cell_step.cubin')) > 10000


def test_oversized_tile_is_refused_before_the_compiler():
    s = workloads.stencil_only(myokit_b200.SimulationCUDA, 512, 256,
                               precision=myokit.DOUBLE_PRECISION)
    s.set_kernel_options(block=(128, 2), cells_per_thread=2, rows_per_thread=8)
    with pytest.raises(ValueError, match='shared memory'):
        s.kernel_source()
    s.set_kernel_options(block=(128, 2), cells_per_thread=2, rows_per_thread=4)
    assert 'mkb_cell_step' in s.kernel_source().code


def test_opt_in_arithmetic_variants_generate_and_build():
    # prepared, off by default: Estrin-scheme exp and side-by-side division
    s = workloads.c3_hetero(myokit_b200.SimulationCUDA, nx=16)
    base = s.kernel_source().code
    assert '#define MKB_DIV_PARALLEL 0' in base
    assert 'mkb_exp_poly(' in base[base.index('extern "C" __global__'):]
    s.set_kernel_options(fast_exp='estrin', div_parallel=True)
    src = s.kernel_source()
    body = src.code[src.code.index('extern "C" __global__'):]
    assert '#define MKB_DIV_PARALLEL 1' in src.code
    assert 'mkb_exp_estrin(' in body and 'mkb_exp_poly(' not in body
    cubin, log = capi.jit_compile(src.code, src.options)
    assert len(cubin) > 10000


def test_command_line(tmp_path, monkeypatch, capsys):
    # python -m myokit_b200 cuda / cuda-select (cf. myokit opencl, opencl-select)
    from myokit_b200 import __main__ as cli
    from myokit_b200 import cuda as cuda_mod
    monkeypatch.setattr(cuda_mod.CUDA, '_path',
                        staticmethod(lambda: str(tmp_path / 'sel.ini')))
    monkeypatch.delenv('MYOKIT_CUDA_DEVICE', raising=False)
    assert cli.main(['cuda']) == 0
    out = capsys.readouterr().out
    assert ('Device 0' in out) if HAS_GPU else ('No CUDA devices found.' in out)
    assert cli.main([]) == 2
    assert cli.main(['cuda-select', '--clear']) == 0
    if HAS_GPU:
        assert cli.main(['cuda-select', '--device', '0']) == 0
        assert cuda_mod.CUDA.load_selection() == 0
        assert cli.main(['cuda-select', '--device', '99']) == 1
    else:
        assert cli.main(['cuda-select', '--device', '0']) == 1
    # two devices, without the hardware: the selection is stored and shown
    info = dict(name='B200', compute_capability=(10, 0), sm_count=148,
                clock_khz=1965000, total_mem=180 << 30, l2_bytes=126 << 20,
                smem_per_block_optin=232448)
    monkeypatch.setattr(capi, 'device_count', lambda: 2)
    monkeypatch.setattr(capi, 'device_info', lambda i: info)
    capsys.readouterr()
    assert cli.main(['cuda-select', '--device', '1']) == 0
    assert cuda_mod.CUDA.load_selection() == 1
    assert cli.main(['cuda']) == 0
    out = capsys.readouterr().out
    assert 'Device 1: B200' in out and '(selected)' in out
    assert 'Multiprocessors   : 148' in out


def test_host_state_stays_one_cell_until_cells_differ():
    # the initial state is the model's, for every cell: no n_cells-long array
    # exists (an 8192 x 8192 grid would need 25.8 GB of host memory for it)
    # until a caller asks for one or changes a single cell
    m, p, _ = myokit.load('example')
    n = m.count_states()
    s = myokit_b200.SimulationCUDA(m, p, ncells=(5, 4), precision=DP)
    assert s._state_full is None and s._default_full is None
    init = list(m.initial_values(True))
    assert s.state(3, 2) == init and s.default_state(0, 0) == init
    assert s._state_full is None                    # single-cell queries do not tile it
    s.set_state([float(k) for k in range(n)])       # one cell's values for all
    assert s._state_full is None and s.state(4, 3) == [float(k) for k in range(n)]
    s.reset()
    assert s._state_full is None and s.state(1, 1) == init
    # one cell differs: now there is an array, in the reference's layout
    s.set_state([9.0] * n, 2, 1)
    assert s._state_full is not None
    full = s.state_array()
    assert full.shape == (5 * 4 * n,)
    assert list(full[(2 + 1 * 5) * n:(3 + 1 * 5) * n]) == [9.0] * n
    assert list(full[:n]) == init
    # the simulation's own array handed back is adopted, not copied
    mine = s.state_array(copy=False)
    mine[0] = -1.0
    s.set_state(mine)
    assert s.state_array(copy=False) is mine and s.state(0, 0)[0] == -1.0
    # a full vector from the caller is copied
    other = np.zeros(5 * 4 * n)
    s.set_state(other)
    other[0] = 7.0
    assert s.state(0, 0)[0] == 0.0
    # default state follows the same rules
    s.set_default_state([1.0] * n)
    assert s._default_full is None and s.default_state(4, 3) == [1.0] * n
    s.set_default_state([2.0] * n, 0, 0)
    assert s.default_state(0, 0) == [2.0] * n and s.default_state(1, 0) == [1.0] * n
    assert len(s.default_state()) == 5 * 4 * n


def test_stage_applies_where_tma_can_describe_the_tile_and_two_blocks_fit():
    from myokit_b200 import kernelgen
    DP, SP = myokit.DOUBLE_PRECISION, myokit.SINGLE_PRECISION
    H = kernelgen.DIFF_HOMOGENEOUS
    assert kernelgen.stage_applies(DP, 48, (128, 2), H, 2048)
    assert kernelgen.stage_applies(DP, 48, (128, 2), kernelgen.DIFF_NONE, 1000)
    assert not kernelgen.stage_applies(DP, 48, (128, 2), H, 2047)          # rows not 16-byte multiples
    assert not kernelgen.stage_applies(SP, 48, (128, 2), H, 2050)
    assert kernelgen.stage_applies(SP, 48, (128, 2), H, 2052)
    assert not kernelgen.stage_applies(DP, 48, (128, 2), kernelgen.DIFF_CONNECTIONS, 2048)
    assert not kernelgen.stage_applies(DP, 48, (128, 2), H, 2048, cells_per_thread=2)
    assert not kernelgen.stage_applies(DP, 48, (128, 2), H, 2048, persistent=True)
    assert not kernelgen.stage_applies(DP, 48, (128, 2), H, 2048, junction='fiber')
    assert not kernelgen.stage_applies(DP, 60, (128, 2), H, 2048)          # 120 KB: two blocks do not fit
    assert kernelgen.stage_applies(DP, 60, (128, 1), H, 2048)
    assert not kernelgen.stage_applies(DP, 48, (4, 2), H, 2048)            # 64-byte tile planes
    # what SimulationCUDA derives from it
    from myokit_b200 import workloads
    s = workloads.c3_hetero(myokit_b200.SimulationCUDA, nx=64, ny=24)
    src = s.kernel_source()
    assert src.kernel_flags & capi.KERNEL_STAGE and src.smem_bytes == 128 + 47 * 64 * 4 * 8
    assert src.plane_stride == 64 * 24 and 'MKB_SHFL_UP(vc, 1)' in src.code
    s.set_kernel_options(stage=False)
    src = s.kernel_source()
    assert not (src.kernel_flags & capi.KERNEL_STAGE) and src.smem_bytes == 0
    assert 'MKB_PREFETCH_L' in src.code         # the plain kernel's own defaults are back


def test_staged_kernel_forms_compile_and_use_the_tma_unit(tmp_path):
    # cross-compiled for sm_100a here (no GPU): the SASS must hold the TMA
    # loads / stores, the arrival-barrier waits, the warp shuffles of the
    # neighbour exchange and, for the tile loop, the asynchronous copies
    import subprocess

    def sass(opts, comm=None, nx=2048, ny=None):
        kw = {} if comm is None else dict(comm=comm)
        s = workloads.c3_hetero(myokit_b200.SimulationCUDA, nx=nx, ny=ny or nx, **kw)
        s.set_kernel_options(**opts)
        src = s.kernel_source()
        cubin, log = capi.jit_compile(src.code, src.options + ('--ptxas-options=-v',))
        path = tmp_path / 'k.cubin'
        path.write_bytes(cubin)
        text = subprocess.check_output(['cuobjdump', '-sass', str(path)]).decode()
        return src, log, text
    # what bench.py runs on one GPU
    src, log, text = sass({})
    assert src.kernel_flags == capi.KERNEL_STAGE and src.plane_stride == 2048 * 2048
    assert 'UTMALDG.3D' in text and 'UTMASTG.3D' in text and 'SYNCS' in text
    assert 'SHFL.UP' in text and 'SHFL.DOWN' in text
    assert 'Used 128 registers' in log or 'Used 127 registers' in log
    m = re.search(r'(\d+) bytes spill stores', log)
    assert int(m.group(1)) <= 16                        # (the plain kernel: 156)
    assert 'CCTL' not in text                            # no software prefetch left
    # small grids: steps overlap, results leave by plain stores
    src, log, text = sass({}, nx=512)
    assert src.kernel_flags == capi.KERNEL_STAGE | capi.KERNEL_OVERLAP
    assert 'UTMALDG.3D' in text and 'UTMASTG' not in text
    # the loop over tiles
    src, log, text = sass(dict(tile_loop=True))
    assert src.kernel_flags & capi.KERNEL_TILE_LOOP and (src.kernel_flags >> 8) == 2
    assert 'LDGSTS' in text and 'UTMALDG.3D' in text and 'UTMASTG.3D' in text
    # two resident blocks fit: dynamic + static shared memory + 1 KiB each
    static = int(re.search(r'(\d+) bytes smem', log).group(1))
    assert 2 * (src.smem_bytes + static + 1024) <= 233472
    # the plain kernel, compiled for its stride
    src, log, text = sass(dict(stage=False))
    assert src.kernel_flags == 0 and 'UTMALDG' not in text and 'CCTL' in text
    assert 'constexpr unsigned long long stride = 4194304ull;' in src.code
    src, log, text = sass(dict(stage=False, plane_stride=False))
    assert src.plane_stride == 0 and 'const unsigned long long stride = g.stride;' in src.code
