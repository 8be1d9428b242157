"""
The generated CUDA source itself, checked on the CPU: ``tests/cuda_shim``
compiles the very text ``SimulationCUDA`` would hand to NVRTC as host C++ (a
thread block = a set of fibers, ``__syncthreads`` = a fiber barrier) and steps
it with the library's own schedule. Compared with the oracle on the same
inputs:

* with the arithmetic rewrites off (``pow`` / IEEE division / libm ``exp``,
  no FMA contraction) the kernel must reproduce the oracle BIT FOR BIT — any
  difference is an indexing, ordering or formula error of the generator;
* with the default rewrites (and the opt-in ones) it must stay within the
  fp64 bar of 1e-6 mV; observed ~2e-12 mV, asserted at 1e-9.

This is test infrastructure: the product has no CPU path.
"""
import numpy as np
import pytest

import myokit_b200
import myokit
from myokit_b200 import workloads
from oracle.oracle import OracleSimulation

import cuda_shim

DP = myokit.DOUBLE_PRECISION
SP = myokit.SINGLE_PRECISION
EXACT = dict(fast_div=False, fast_exp=False, pow_multiply=False, fmad=False,
             const_div=False, fast_libm=False)


def oracle_fields(o, duration, li, nx, ny, names):
    log, state = o.run(duration, log=['engine.time'] + names, log_interval=li)
    nt = len(log['engine.time'])
    out = {'time': np.array(log['engine.time'])}
    for name in names:
        if ny > 1:
            out[name] = np.array([[log['%d.%d.%s' % (x, y, name)][k]
                                   for y in range(ny) for x in range(nx)]
                                  for k in range(nt)])
        else:
            out[name] = np.array([[log['%d.%s' % (x, name)][k]
                                   for x in range(nx)] for k in range(nt)])
    return out, np.asarray(state)


def both(make, options, duration, li, nx, ny, inter_log=()):
    a = make(myokit_b200.SimulationCUDA)
    a.set_kernel_options(**options)
    got = cuda_shim.run_on_host(a, duration, log_interval=li,
                                inter_log=inter_log)
    names = ['membrane.V', 'membrane.i_diff'] if a._diffusion_enabled \
        else ['membrane.V']
    want, wstate = oracle_fields(make(OracleSimulation), duration, li, nx, ny,
                                 names + list(inter_log))
    assert np.array_equal(got['time'], want['time'])
    return got, want, wstate


def lr91_2d(cls, precision=DP, nx=10, ny=7):
    m, _, _ = myokit.load('example')
    p = myokit.pacing.blocktrain(period=1000, duration=2, offset=1)
    s = cls(m, p, ncells=(nx, ny), precision=precision)
    s.set_conductance(9, 6)
    s.set_paced_cells(3, ny, 0, 0)
    s.set_step_size(0.005)
    return s


def test_lr91_2d_exact_options_bit_for_bit():
    # ragged grid: 10 x 7 cells under 8 x 4 thread blocks
    got, want, wstate = both(lr91_2d, dict(EXACT, block=(8, 4)), 5.0, 0.5, 10, 7,
                             inter_log=['ina.INa'])
    assert want['membrane.V'].max() > 0         # the paced edge fired
    assert np.array_equal(got['V'], want['membrane.V'])
    assert np.array_equal(got['idiff'], want['membrane.i_diff'])
    assert np.array_equal(got['inter'][:, 0], want['ina.INa'])
    assert np.array_equal(got['state'].ravel(), wstate)


def test_decker_hetero_rush_larsen_fields():
    def make(cls):
        return workloads.c3_hetero(cls, nx=12, ny=9)
    got, want, wstate = both(make, dict(EXACT, block=(8, 4)), 3.0, 0.5, 12, 9)
    assert want['membrane.V'].max() > 0
    assert np.array_equal(got['V'], want['membrane.V'])
    assert np.array_equal(got['idiff'], want['membrane.i_diff'])
    assert np.array_equal(got['state'].ravel(), wstate)


@pytest.mark.parametrize('options', [
    dict(),                                     # what bench.py runs
    dict(div_parallel=True),
    dict(fast_exp='estrin'),
    dict(fast_exp='table'),
    dict(const_div=False),
    dict(load_ahead=4, lazy_state=True),
    dict(lazy_state=False),
    dict(div_cubic=True),
    dict(fast_exp='stab'),
    dict(div_cubic=True, fast_exp='stab', prefetch='l1', load_ahead=8),
], ids=['default', 'div_parallel', 'estrin', 'table', 'no_const_div',
        'load_ahead4', 'eager_loads', 'div_cubic', 'exp_stab',
        'cubic_stab_prefetch'])
def test_decker_arithmetic_rewrites_within_the_fp64_bar(options):
    def make(cls):
        return workloads.c3_hetero(cls, nx=12, ny=9)
    got, want, wstate = both(make, dict(options, block=(8, 4)), 3.0, 0.5, 12, 9)
    assert np.abs(got['V'] - want['membrane.V']).max() <= 1e-9
    assert np.abs(got['idiff'] - want['membrane.i_diff']).max() <= 1e-9
    rel = np.abs(got['state'].ravel() - wstate) / (np.abs(wstate) + 1e-12)
    assert rel.max() <= 1e-6


@pytest.mark.parametrize('cpt,rpt,block', [(2, 1, (4, 2)), (2, 4, (4, 2)),
                                           (4, 2, (2, 2))])
def test_register_patch_path_equals_oracle_fp64(cpt, rpt, block):
    # several cells per thread, rim exchange through shared memory; grid
    # sizes that leave partial patches at the right and bottom edges
    def make(cls):
        return lr91_2d(cls, nx=20, ny=11)
    opts = dict(EXACT, block=block, cells_per_thread=cpt, rows_per_thread=rpt)
    got, want, wstate = both(make, opts, 4.0, 0.5, 20, 11)
    assert want['membrane.V'].max() > 0
    assert np.array_equal(got['V'], want['membrane.V'])
    assert np.array_equal(got['idiff'], want['membrane.i_diff'])
    assert np.array_equal(got['state'].ravel(), wstate)


@pytest.mark.parametrize('hetero', [False, True])
def test_stencil_only_fp32_patch_path(hetero):
    def make(cls):
        s = workloads.stencil_only(cls, 24, 10, precision=SP, hetero=hetero)
        # a bump to diffuse
        st = np.array(s.state(), dtype=float).reshape(10, 24)
        st[3:6, 5:9] = 20.0
        s.set_state(list(st.ravel()))
        return s
    opts = dict(fmad=False, block=(2, 2), cells_per_thread=4, rows_per_thread=2)
    got, want, wstate = both(make, opts, 2.0, 0.25, 24, 10)
    assert np.array_equal(got['V'], want['membrane.V'])
    assert np.array_equal(got['state'].ravel(), wstate)


def test_cable_paced_list_and_connections():
    m, _, _ = myokit.load('example')
    p = myokit.pacing.blocktrain(period=1000, duration=2, offset=1)
    n = 37

    def cable(cls):
        s = cls(m, p, ncells=n, precision=DP)
        s.set_conductance(8)
        s.set_paced_cell_list([0, 1, 5, 36])
        return s

    def graph(cls):
        s = cls(m, p, ncells=n, precision=DP)
        edges = [(i, i + 1, 8.0) for i in range(n - 1)] + [(3, 30, 0.5), (0, 36, 1.5)]
        s.set_connections(edges)
        s.set_paced_cell_list([0, 1, 5, 36])
        return s
    for make in (cable, graph):
        got, want, wstate = both(make, dict(EXACT, block=(16, 1)), 4.0, 0.5, n, 1)
        assert want['membrane.V'].max() > 0
        assert np.array_equal(got['V'], want['membrane.V'])
        assert np.array_equal(got['idiff'], want['membrane.i_diff'])
        assert np.array_equal(got['state'].ravel(), wstate)


def test_uncoupled_cells_with_a_field():
    m, _, _ = myokit.load('example')
    p = myokit.pacing.blocktrain(period=1000, duration=2, offset=1)
    n = 21
    gna = 16.0 * (1 + 0.3 * np.sin(np.arange(n)))

    def make(cls):
        s = cls(m, p, ncells=n, diffusion=False, precision=DP)
        s.set_field('ina.gNa', gna)
        return s
    got, want, wstate = both(make, dict(EXACT, block=(8, 1)), 4.0, 0.5, n, 1)
    assert want['membrane.V'].max() > 0
    assert np.array_equal(got['V'], want['membrane.V'])
    assert np.array_equal(got['state'].ravel(), wstate)


def test_lr91_fp32_default_options():
    def make(cls):
        return lr91_2d(cls, precision=SP)
    got, want, wstate = both(make, dict(block=(8, 4)), 5.0, 0.5, 10, 7)
    assert got['real_size'] == 4
    # fp32 with fast division / FMA contraction: close, not identical
    assert np.abs(got['V'] - want['membrane.V']).max() <= 5e-2


# ---------------------------------------------------------------------------
# Thread order: nothing may depend on the order in which the threads of a block
# reach a barrier. The shim can run them first-to-last or last-to-first.
# (Found this way: threads beyond the last row used to write their — unused —
# centre value into the tile slot that holds the halo of the row below them;
# harmless on one GPU, a write-write race with the ghost row in a row slab
# whose height is not a multiple of the block height.)
# ---------------------------------------------------------------------------
def test_results_do_not_depend_on_thread_order():
    def make(cls):
        return lr91_2d(cls, nx=10, ny=7)
    runs = []
    for reverse in (False, True):
        a = make(myokit_b200.SimulationCUDA)
        a.set_kernel_options(**dict(EXACT, block=(8, 4)))
        runs.append(cuda_shim.run_on_host(a, 3.0, log_interval=0.5, reverse=reverse))
    assert np.array_equal(runs[0]['V'], runs[1]['V'])
    assert np.array_equal(runs[0]['state'], runs[1]['state'])

    def make(cls):
        return lr91_2d(cls, nx=20, ny=11)
    runs = []
    for reverse in (False, True):
        a = make(myokit_b200.SimulationCUDA)
        a.set_kernel_options(**dict(EXACT, block=(4, 2), cells_per_thread=2,
                                    rows_per_thread=4))
        runs.append(cuda_shim.run_on_host(a, 3.0, log_interval=0.5, reverse=reverse))
    assert runs[0]['V'].max() > 0
    assert np.array_equal(runs[0]['V'], runs[1]['V'])
    assert np.array_equal(runs[0]['state'], runs[1]['state'])


@pytest.mark.parametrize('lean', [False, True], ids=['slab', 'slab_lean'])
@pytest.mark.parametrize('nslab', [2, 3])
def test_row_slab_kernels_equal_the_whole_grid(nslab, lean):
    # 11 rows: slab heights 5/6 or 3/4/4 under 4-row thread blocks, so every
    # slab has a partial last block row; conductance fields cross the cuts
    def make(comm):
        kw = {} if comm is None else dict(comm=comm)
        return workloads.c3_hetero(myokit_b200.SimulationCUDA, nx=12, ny=11, **kw)
    opts = dict(EXACT, block=(8, 4))
    whole = make(None)
    whole.set_kernel_options(**opts)
    one = cuda_shim.run_on_host(whole, 3.0, log_interval=0.5)
    assert one['V'].max() > 0
    for reverse in (False, True):
        out = cuda_shim.run_slabs_on_host(make, nslab, 3.0, 0.5,
                                          dict(opts, slab_lean=lean), reverse=reverse)
        assert out['halo_error'] == 0
        assert np.array_equal(out['V'], one['V'])
        assert np.array_equal(out['idiff'], one['idiff'])
        assert np.array_equal(out['state'], one['state'])


@pytest.mark.parametrize('nparts', [2, 3])
def test_partitioned_graph_kernels_equal_the_whole_graph(nparts):
    # a small fibre mesh with long-range edges, cut into contiguous id blocks
    # by the product's own _partition_graph; ghost planes and flags on the host
    m, _, _ = myokit.load('example')
    p = myokit.pacing.blocktrain(period=1000, duration=2, offset=1)
    n, edges = workloads.fibre_mesh(6, 4, 5, extra=0.1, seed=3)

    def make(comm):
        kw = {} if comm is None else dict(comm=comm)
        s = myokit_b200.SimulationCUDA(m, p, ncells=n, precision=DP, **kw)
        s.set_connections(edges)
        s.set_paced_cells(12)
        return s
    opts = dict(EXACT, block=(16, 1))
    whole = make(None)
    whole.set_kernel_options(**opts)
    one = cuda_shim.run_on_host(whole, 5.0, log_interval=0.5)
    assert one['V'].max() > 0
    for reverse in (False, True):
        out = cuda_shim.run_parts_on_host(make, nparts, 5.0, 0.5, opts,
                                          reverse=reverse)
        assert out['halo_error'] == 0
        assert np.array_equal(out['V'], one['V'])
        assert np.array_equal(out['idiff'], one['idiff'])
        assert np.array_equal(out['state'], one['state'])


# ---------------------------------------------------------------------------
# Fibre-tissue junction: two kernels, two models, stepped
# in lockstep, against oracle/fiber_tissue.py (itself pinned against the
# reference's rendered kernels in tests/test_oracle_fiber_tissue.py).
# ---------------------------------------------------------------------------
@pytest.mark.parametrize('options,exact', [(EXACT, True), ({}, False)],
                         ids=['exact', 'default'])
def test_fiber_tissue_junction_kernels(options, exact):
    import os
    from oracle.fiber_tissue import OracleFiberTissue
    data = os.path.join(os.path.dirname(myokit.__file__), 'tests', 'data')
    mf = myokit.load_model(os.path.join(data, 'dn-1985-normalised.mmt'))
    mt = myokit.load_model(os.path.join(data, 'lr-1991.mmt'))
    p = myokit.pacing.blocktrain(1000, 2.0, offset=.01)
    f = myokit_b200.SimulationCUDA(mf, p, ncells=(8, 4), precision=DP)
    f.set_conductance(235, 100)
    f.set_paced_cells(4, 4, 0, 0)
    f.set_step_size(0.0012)
    f.set_kernel_options(block=(8, 2), junction='fiber', **options)
    t = myokit_b200.SimulationCUDA(mt, p, ncells=(8, 6), precision=DP)
    t.set_conductance(9, 5)
    t.set_paced_cells(0, 0, 0, 0)
    t.set_step_size(0.0012)
    t.set_kernel_options(block=(8, 2), junction='tissue', **options)
    out = cuda_shim.run_pair_on_host(f, t, 9.0, 1, 3.0, 0.25)
    o = OracleFiberTissue(
        mf, mt, p, ncells_fiber=(8, 4), ncells_tissue=(8, 6), nx_paced=4,
        g_fiber=(235, 100), g_tissue=(9, 5), g_fiber_tissue=9, dt=0.0012,
        precision=DP)
    names = ['membrane.V', 'membrane.i_diff']
    tt, of, ot = o.run(3.0, names, names, 0.25)
    assert np.array_equal(out['time'], tt)
    assert ot['membrane.V'].max() > 0       # driven through the junction
    if exact:
        assert np.array_equal(out['fiber']['V'], of['membrane.V'])
        assert np.array_equal(out['fiber']['idiff'], of['membrane.i_diff'])
        assert np.array_equal(out['tissue']['V'], ot['membrane.V'])
        assert np.array_equal(out['tissue']['idiff'], ot['membrane.i_diff'])
        assert np.array_equal(out['fiber']['state'], o.fiber_state())
        assert np.array_equal(out['tissue']['state'], o.tissue_state())
    else:
        assert np.abs(out['fiber']['V'] - of['membrane.V']).max() <= 1e-8
        assert np.abs(out['tissue']['V'] - ot['membrane.V']).max() <= 1e-8


def test_junction_needs_a_plain_homogeneous_kernel():
    def make(cls):
        return lr91_2d(cls)
    s = make(myokit_b200.SimulationCUDA)
    s.set_kernel_options(junction='sideways')
    with pytest.raises(ValueError):
        s.kernel_source()
    s = workloads.c3_hetero(myokit_b200.SimulationCUDA, nx=8)
    s.set_kernel_options(junction='fiber')
    with pytest.raises(ValueError):
        s.kernel_source()


def test_lr91_fp32_lean_exp_variant():
    # opt-in fast_exp='ex2' (single precision): same accuracy class as the
    # default fp32 kernel relative to the fp32 oracle
    def make(cls):
        return lr91_2d(cls, precision=SP)
    base, want, _ = both(make, dict(block=(8, 4)), 5.0, 0.5, 10, 7)
    got, want, _ = both(make, dict(block=(8, 4), fast_exp='ex2'), 5.0, 0.5, 10, 7)
    e0 = np.abs(base['V'] - want['membrane.V']).max()
    e1 = np.abs(got['V'] - want['membrane.V']).max()
    assert e1 <= max(5e-2, 3 * e0)
    s = make(myokit_b200.SimulationCUDA)
    s.set_kernel_options(fast_exp='ex2')
    assert 'mkb_expf_ex2(' in s.kernel_source().code.split('extern "C" __global__')[1]
    d = lr91_2d(myokit_b200.SimulationCUDA, precision=DP)
    d.set_kernel_options(fast_exp='ex2')
    with pytest.raises(ValueError):
        d.kernel_source()


# ---------------------------------------------------------------------------
# More models and option combinations, all bit for bit with the rewrites off
# ---------------------------------------------------------------------------
def _data_model(name):
    import os
    return myokit.load_model(os.path.join(
        os.path.dirname(myokit.__file__), 'tests', 'data', name))


@pytest.mark.parametrize('case', ['lr91_rl', 'br77', 'decker_fe', 'lr91_hetero_list_field'])
def test_more_models_bit_for_bit(case):
    p = myokit.pacing.blocktrain(period=1000, duration=2, offset=1)
    nx, ny = 9, 6
    rng = np.random.default_rng(5)
    gxf = rng.uniform(2, 9, size=(ny, nx - 1))
    gyf = rng.uniform(2, 9, size=(ny - 1, nx))
    gna = 16.0 * (1 + 0.2 * rng.uniform(-1, 1, size=(ny, nx)))

    def make(cls):
        if case == 'lr91_rl':
            m, _, _ = myokit.load('example')
            s = cls(m, p, ncells=(nx, ny), precision=DP, rl=True)
            s.set_conductance(8, 5)
            s.set_paced_cells(2, ny, 0, 0)
        elif case == 'br77':
            s = cls(_data_model('beeler-1977-model.mmt'), p, ncells=(nx, ny), precision=DP)
            s.set_conductance(8, 5)
            s.set_paced_cells(2, ny, 0, 0)
        elif case == 'decker_fe':
            s = cls(_data_model('decker-2009.mmt'), p, ncells=(nx, ny), precision=DP)
            s.set_conductance(8, 5)
            s.set_paced_cells(-2, ny, 0, 0)      # the two right-most columns
        else:
            m, _, _ = myokit.load('example')
            s = cls(m, p, ncells=(nx, ny), precision=DP)
            s.set_conductance_field(gxf, gyf)
            s.set_paced_cell_list([(0, 0), (1, 0), (0, 1), (8, 5), (4, 3)])
            s.set_field('ina.gNa', gna)
        s.set_step_size(0.005)
        return s
    got, want, wstate = both(make, dict(EXACT, block=(8, 4)), 3.0, 0.5, nx, ny)
    assert want['membrane.V'].max() > want['membrane.V'][0].max() + 5     # something happened
    assert np.array_equal(got['V'], want['membrane.V'])
    assert np.array_equal(got['idiff'], want['membrane.i_diff'])
    assert np.array_equal(got['state'].ravel(), wstate)


# ---------------------------------------------------------------------------
# Randomised shapes (fixed seeds): grid sizes down to 1 x 1, 1-d and 2-d, block
# shapes, register patches, conductance fields, pacing rectangles anywhere,
# both thread orders; row slabs with 2-4 ranks. All bit for bit.
# ---------------------------------------------------------------------------
@pytest.mark.parametrize('seed', [11, 12])
def test_random_shapes_bit_for_bit(seed):
    m, _, _ = myokit.load('example')
    p = myokit.pacing.blocktrain(period=1000, duration=1, offset=0.2)
    rng = np.random.default_rng(seed)
    for trial in range(5):
        two_d = rng.random() < 0.8
        nx = int(rng.integers(1, 14))
        ny = int(rng.integers(1, 9)) if two_d else 1
        cpt = int(rng.choice([1, 1, 2, 4]))
        rpt = int(rng.choice([1, 2, 3, 4])) if cpt > 1 else 1
        if cpt > 1:
            nx = max(cpt, nx - nx % cpt)
        bx = int(rng.choice([2, 4, 8]))
        by = int(rng.choice([1, 2, 4])) if two_d else 1
        hetero = two_d and nx > 1 and ny > 1 and rng.random() < 0.4
        px, py = int(rng.integers(0, nx)), int(rng.integers(0, ny))
        pnx, pny = int(rng.integers(1, 4)), int(rng.integers(1, 4))
        gxf = rng.uniform(1, 9, size=(ny, max(nx - 1, 0)))
        gyf = rng.uniform(1, 9, size=(max(ny - 1, 0), nx))

        def make(cls):
            s = cls(m, p, ncells=(nx, ny) if two_d else nx, precision=DP)
            if hetero:
                s.set_conductance_field(gxf, gyf)
            elif two_d:
                s.set_conductance(7, 4)
            else:
                s.set_conductance(7)
            if two_d:
                s.set_paced_cells(pnx, pny, px, py)
            else:
                s.set_paced_cells(pnx, x=px)
            return s
        a = make(myokit_b200.SimulationCUDA)
        a.set_kernel_options(block=(bx, by), cells_per_thread=cpt,
                             rows_per_thread=rpt, **EXACT)
        out = cuda_shim.run_on_host(a, 1.5, log_interval=0.25,
                                    reverse=bool(trial & 1))
        log, ostate = make(OracleSimulation).run(
            1.5, log=['engine.time', 'membrane.V'], log_interval=0.25)
        what = (seed, trial, nx, ny, cpt, rpt, bx, by, hetero)
        assert np.array_equal(out['state'].ravel(), np.asarray(ostate)), what


@pytest.mark.parametrize('seed', [21])
def test_random_row_slabs_bit_for_bit(seed):
    m, _, _ = myokit.load('example')
    p = myokit.pacing.blocktrain(period=1000, duration=1, offset=0.2)
    rng = np.random.default_rng(seed)
    for trial in range(4):
        nslab = int(rng.integers(2, 5))
        nx = int(rng.integers(2, 14))
        ny = int(rng.integers(nslab, nslab * 4 + 3))
        bx = int(rng.choice([2, 4, 8]))
        by = int(rng.choice([1, 2, 4]))
        hetero = rng.random() < 0.5
        lean = bool(rng.random() < 0.5)
        px, py = int(rng.integers(0, nx)), int(rng.integers(0, ny))
        pnx, pny = int(rng.integers(1, 4)), int(rng.integers(1, ny + 1))
        gxf = rng.uniform(1, 9, size=(ny, nx - 1))
        gyf = rng.uniform(1, 9, size=(ny - 1, nx))

        def make(comm):
            kw = {} if comm is None else dict(comm=comm)
            s = myokit_b200.SimulationCUDA(m, p, ncells=(nx, ny),
                                           precision=DP, **kw)
            if hetero:
                s.set_conductance_field(gxf, gyf)
            else:
                s.set_conductance(7, 4)
            s.set_paced_cells(pnx, pny, px, py)
            return s
        opts = dict(EXACT, block=(bx, by))
        whole = make(None)
        whole.set_kernel_options(**opts)
        one = cuda_shim.run_on_host(whole, 1.5, log_interval=0.25)
        out = cuda_shim.run_slabs_on_host(
            make, nslab, 1.5, 0.25, dict(opts, slab_lean=lean),
            reverse=bool(trial & 1))
        what = (seed, trial, nslab, nx, ny, bx, by, hetero, lean)
        assert out['halo_error'] == 0, what
        assert np.array_equal(out['state'], one['state']), what
        assert np.array_equal(out['V'], one['V']), what


def test_generated_kernel_equals_reference_simulation1d_golden():
    # The bench model, straight against what the REFERENCE computed
    # (tests/golden/sim1d_decker_rl.npz, made by myokit.Simulation1d): with
    # the rewrites off the generated CUDA source gives the same bits; with
    # the defaults it stays within 1e-9 mV.
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden',
                             'sim1d_decker_rl.npz'))
    m = _data_model('decker-2009.mmt')
    p = myokit.pacing.blocktrain(duration=2, offset=1, period=1000)
    want_v = np.array([g['log:%d.membrane.V' % x] for x in range(12)]).T
    want_i = np.array([g['log:%d.membrane.i_diff' % x] for x in range(12)]).T
    for options, exact in ((EXACT, True), ({}, False)):
        s = myokit_b200.SimulationCUDA(m, p, ncells=12, precision=DP, rl=True)
        s.set_conductance(10)
        s.set_paced_cells(3)
        s.set_step_size(0.005)
        s.set_kernel_options(block=(8, 1), **options)
        out = cuda_shim.run_on_host(s, 12, log_interval=0.5)
        assert np.array_equal(out['time'], g['log:engine.time'])
        if exact:
            assert np.array_equal(out['V'], want_v)
            assert np.array_equal(out['idiff'], want_i)
            assert np.array_equal(out['state'].ravel(), g['state'])
        else:
            assert np.abs(out['V'] - want_v).max() <= 1e-9
            rel = np.abs(out['state'].ravel() - g['state']) / (np.abs(g['state']) + 1e-12)
            assert rel.max() <= 1e-6


# ---------------------------------------------------------------------------
# Persistent kernel (opt-in): the block is the grid, all steps up to the next
# logged one inside one launch, states in registers in between.
# ---------------------------------------------------------------------------
def test_persistent_kernel_cable_and_grid_bit_for_bit():
    m, _, _ = myokit.load('example')
    p = myokit.pacing.blocktrain(period=1000, duration=0.5, offset=1)

    def cable(cls):
        s = cls(m, p, ncells=40, precision=DP)
        s.set_conductance(10)
        s.set_paced_cells(5)
        return s
    got, want, wstate = both(cable, dict(EXACT, persistent=True), 6.0, 0.5, 40, 1,
                             inter_log=['ina.INa'])
    assert want['membrane.V'].max() > 0
    assert np.array_equal(got['V'], want['membrane.V'])
    assert np.array_equal(got['idiff'], want['membrane.i_diff'])
    assert np.array_equal(got['inter'][:, 0], want['ina.INa'])
    assert np.array_equal(got['state'].ravel(), wstate)

    rng = np.random.default_rng(2)
    gxf = rng.uniform(2, 9, size=(6, 9))
    gyf = rng.uniform(2, 9, size=(5, 10))
    gna = 16 * (1 + 0.1 * rng.uniform(-1, 1, size=(6, 10)))
    p2 = myokit.pacing.blocktrain(period=1000, duration=2, offset=1)

    def grid(cls):
        s = cls(m, p2, ncells=(10, 6), precision=DP, rl=True)
        s.set_conductance_field(gxf, gyf)
        s.set_paced_cells(3, 6, 0, 0)
        s.set_field('ina.gNa', gna)
        return s
    for reverse_opts in (dict(), dict(block=(16, 8))):
        got, want, wstate = both(grid, dict(EXACT, persistent=True, **reverse_opts),
                                 4.0, 0.5, 10, 6)
        assert np.array_equal(got['V'], want['membrane.V'])
        assert np.array_equal(got['idiff'], want['membrane.i_diff'])
        assert np.array_equal(got['state'].ravel(), wstate)

    def cells(cls):
        s = cls(m, p2, ncells=21, diffusion=False, precision=DP)
        s.set_field('ina.gNa', 16.0 * (1 + 0.3 * np.sin(np.arange(21))))
        return s
    got, want, wstate = both(cells, dict(EXACT, persistent=True), 4.0, 0.5, 21, 1)
    assert np.array_equal(got['V'], want['membrane.V'])
    assert np.array_equal(got['state'].ravel(), wstate)


def test_persistent_kernel_limits_and_name():
    m, _, _ = myokit.load('example')
    s = myokit_b200.SimulationCUDA(m, None, ncells=128, precision=DP)
    s.set_kernel_options(persistent=True)
    src = s.kernel_source()
    assert src.kernel_name == 'mkb_cell_step_persistent' and src.persistent
    assert src.block == (128, 1)
    assert 'mkb_cell_step_persistent(const MkbGridArgs g' in src.code
    from myokit_b200 import capi
    cubin, log = capi.jit_compile(src.code, src.options)
    assert len(cubin) > 10000
    big = myokit_b200.SimulationCUDA(m, None, ncells=(64, 32), precision=DP)
    big.set_kernel_options(persistent=True)
    with pytest.raises(ValueError, match='fits one thread block'):
        big.kernel_source()
    g = myokit_b200.SimulationCUDA(m, None, ncells=16, precision=DP)
    g.set_connections([(0, 1, 1.0)])
    g.set_kernel_options(persistent=True)
    with pytest.raises(ValueError):
        g.kernel_source()
    # grids of up to 128 cells take it by default, larger ones the one-step kernel
    d = myokit_b200.SimulationCUDA(m, None, ncells=128, precision=DP)
    assert d.kernel_source().kernel_name == 'mkb_cell_step_persistent'
    d = myokit_b200.SimulationCUDA(m, None, ncells=129, precision=DP)
    assert d.kernel_source().kernel_name == 'mkb_cell_step'


# ---------------------------------------------------------------------------
# split_gates (opt-in): gating variables with voltage-dependent rates in a
# kernel of their own, launched after the big one.
# ---------------------------------------------------------------------------
def test_split_gates_two_kernels_equal_one():
    def make(cls):
        return workloads.c3_hetero(cls, nx=12, ny=9)
    s = make(myokit_b200.SimulationCUDA)
    s.set_kernel_options(split_gates=True)
    src = s.kernel_source()
    assert src.gate_kernel and 'ina.m' in src.gate_states and 'ikr.a' in src.gate_states
    assert 'membrane.V' not in src.gate_states and len(src.gate_states) == 10
    big, gates = src.code.split('mkb_gate_step(const MkbGridArgs g')
    # the big kernel neither computes the gates' rates nor stores the gates
    assert 'V_ina_h_alpha' not in big.split('extern "C" __global__')[1]
    assert 'V_ina_h_alpha' in gates
    k = s._model.get('ina.m').index()
    assert 'MKB_AT(state_c, %d) =' % k not in big
    assert 'MKB_AT(state_c, %d) =' % k in gates
    # bit for bit with the rewrites off (a logged intermediary of each kernel)
    got, want, wstate = both(make, dict(EXACT, block=(8, 4), split_gates=True),
                             3.0, 0.5, 12, 9, inter_log=['ina.h.alpha', 'ikr.IKr'])
    assert want['membrane.V'].max() > 0
    assert np.array_equal(got['V'], want['membrane.V'])
    assert np.array_equal(got['idiff'], want['membrane.i_diff'])
    assert np.array_equal(got['inter'][:, 0], want['ina.h.alpha'])
    assert np.array_equal(got['inter'][:, 1], want['ikr.IKr'])
    assert np.array_equal(got['state'].ravel(), wstate)
    # defaults: as close to the oracle as the single kernel
    got, want, wstate = both(make, dict(block=(8, 4), split_gates=True), 3.0, 0.5, 12, 9)
    assert np.abs(got['V'] - want['membrane.V']).max() <= 1e-9
    rel = np.abs(got['state'].ravel() - wstate) / (np.abs(wstate) + 1e-12)
    assert rel.max() <= 1e-6


def test_split_gates_other_models_and_no_gates():
    # LR1991 forward Euler: m, h, j, d, f, x depend on V only
    def make(cls):
        return lr91_2d(cls, nx=9, ny=6)
    s = make(myokit_b200.SimulationCUDA)
    s.set_kernel_options(split_gates=True)
    src = s.kernel_source()
    assert src.gate_kernel and set(src.gate_states) >= {'ina.m', 'ina.h', 'ina.j'}
    got, want, wstate = both(make, dict(EXACT, block=(8, 4), split_gates=True), 4.0, 0.5, 9, 6)
    assert np.array_equal(got['V'], want['membrane.V'])
    assert np.array_equal(got['state'].ravel(), wstate)
    # uncoupled cells, Rush-Larsen
    m, _, _ = myokit.load('example')
    p = myokit.pacing.blocktrain(period=1000, duration=2, offset=1)

    def cells(cls):
        return cls(m, p, ncells=13, diffusion=False, precision=DP, rl=True)
    c = cells(myokit_b200.SimulationCUDA)
    c.set_kernel_options(split_gates=True)
    assert not c.kernel_source().gate_kernel    # V is not double-buffered here
    got, want, wstate = both(cells, dict(EXACT, block=(8, 1), split_gates=True), 4.0, 0.5, 13, 1)
    assert np.array_equal(got['state'].ravel(), wstate)
    # a model without such states keeps its single kernel
    st = workloads.stencil_only(myokit_b200.SimulationCUDA, 16, 8, precision=DP)
    st.set_kernel_options(split_gates=True)
    assert not st.kernel_source().gate_kernel


def test_split_gates_with_row_slabs():
    def make(comm):
        kw = {} if comm is None else dict(comm=comm)
        return workloads.c3_hetero(myokit_b200.SimulationCUDA, nx=12, ny=11, **kw)
    opts = dict(EXACT, block=(8, 4))
    whole = make(None)
    whole.set_kernel_options(**opts)
    one = cuda_shim.run_on_host(whole, 3.0, log_interval=0.5)
    out = cuda_shim.run_slabs_on_host(make, 3, 3.0, 0.5, dict(opts, split_gates=True))
    assert out['halo_error'] == 0
    assert np.array_equal(out['V'], one['V'])
    assert np.array_equal(out['state'], one['state'])


@pytest.mark.parametrize('precision,nx,ny', [(SP, 140, 37), (DP, 70, 41)])
@pytest.mark.parametrize('hetero', [False, True])
def test_streaming_kernel_tma_tiles_equal_oracle(precision, nx, ny, hetero):
    # kernelgen stream=True: persistent blocks walk tiles; the V tile + halo
    # arrives by TMA (here: the shim's synchronous box copy with zero fill),
    # left / right neighbours by warp shuffle. Ragged grids, 3 and 5 resident
    # blocks, both thread orders: the bits of the oracle.
    from myokit_b200 import workloads

    def make(cls):
        return workloads.stencil_only(cls, nx, ny, precision=precision, hetero=hetero)
    a = make(myokit_b200.SimulationCUDA)
    a.set_kernel_options(stream=True, fmad=False)
    src = a.kernel_source()
    assert src.kernel_flags & 2 and src.block == ((32, 4) if precision == SP else (32, 16))
    assert 'MKB_TMA_LOAD_2D' in src.code and '__grid_constant__' in src.code
    got = cuda_shim.run_on_host(a, 2.5, log_interval=0.5)
    rev = cuda_shim.run_on_host(a, 2.5, log_interval=0.5, reverse=True, stream_blocks=5)
    want, wstate = oracle_fields(make(OracleSimulation), 2.5, 0.5, nx, ny,
                                 ['membrane.V', 'membrane.i_diff'])
    assert want['membrane.V'].max() > -72       # the paced edge moved
    assert np.array_equal(got['V'], want['membrane.V'])
    assert np.array_equal(rev['V'], want['membrane.V'])
    assert np.array_equal(got['idiff'], want['membrane.i_diff'])
    assert np.array_equal(got['state'].ravel(), wstate)


def test_streaming_kernel_falls_back_where_tma_cannot_describe_the_grid():
    from myokit_b200 import workloads
    # rows that are not 16-byte multiples, 1-d cables: the register-patch path
    for nx, ny in ((141, 37), (64, 1)):
        s = workloads.stencil_only(myokit_b200.SimulationCUDA, nx, ny if ny > 1 else None,
                                   precision=SP) if ny > 1 else None
        if s is None:
            continue
        s.set_kernel_options(stream=True)
        assert not (s.kernel_source().kernel_flags & 2)


# ---------------------------------------------------------------------------
# Staged states (option stage): every state plane's tile through shared memory
# by TMA — on the host the shim's synchronous box copies — must not change a bit
# ---------------------------------------------------------------------------
@pytest.mark.parametrize('block,nx,ny', [((8, 4), 12, 9), ((16, 2), 16, 4), ((8, 2), 10, 5)])
def test_staged_states_equal_oracle_bit_for_bit(block, nx, ny):
    def make(cls):
        return workloads.c3_hetero(cls, nx=nx, ny=ny)
    opts = dict(EXACT, block=block, stage=True, load_ahead=2, stage_group=5)
    a = make(myokit_b200.SimulationCUDA)
    a.set_kernel_options(**opts)
    src = a.kernel_source()
    assert src.kernel_flags & 8 and src.smem_bytes > 0
    assert 'MKB_TMA_LOAD_3D' in src.code and 'MKB_PREFETCH_L1(&' not in src.code
    got, want, wstate = both(make, opts, 3.0, 0.5, nx, ny)
    assert want['membrane.V'].max() > 0
    assert np.array_equal(got['V'], want['membrane.V'])
    assert np.array_equal(got['idiff'], want['membrane.i_diff'])
    assert np.array_equal(got['state'].ravel(), wstate)
    # last-to-first thread order: the same bits
    b = make(myokit_b200.SimulationCUDA)
    b.set_kernel_options(**opts)
    rev = cuda_shim.run_on_host(b, 3.0, log_interval=0.5, reverse=True)
    assert np.array_equal(rev['state'], got['state'])


def test_staged_states_uncoupled_cells_and_default_arithmetic():
    m, p, _ = myokit.load('example')

    def make(cls):
        s = cls(m, p, ncells=24, diffusion=False, precision=DP, rl=True)
        s.set_field('ina.gNa', np.linspace(8, 16, 24))
        return s
    opts = dict(EXACT, block=(16, 1), stage=True)
    a = make(myokit_b200.SimulationCUDA)
    a.set_kernel_options(**opts)
    assert a.kernel_source().kernel_flags & 8
    got = cuda_shim.run_on_host(a, 2.0, log_interval=0.5)
    log, ostate = make(OracleSimulation).run(2.0, log=['engine.time'], log_interval=0.5)
    assert np.array_equal(got['state'].ravel(), np.asarray(ostate))
    # the rewrites the bench runs with (in-line division, exp, libm): the fp64 bar
    def make3(cls):
        return workloads.c3_hetero(cls, nx=12, ny=9)
    got, want, wstate = both(make3, dict(block=(8, 4), stage=True), 3.0, 0.5, 12, 9)
    assert np.max(np.abs(got['V'] - want['membrane.V'])) < 1e-9


def test_staged_states_fall_back_where_tma_cannot_describe_the_rows():
    # rows of 11 doubles are not 16-byte multiples; connection graphs and the
    # register-patch path do not stage
    s = workloads.c3_hetero(myokit_b200.SimulationCUDA, nx=11, ny=4)
    s.set_kernel_options(stage=True, block=(8, 4))
    src = s.kernel_source()
    assert not (src.kernel_flags & 8) and src.smem_bytes == 0
    m, p, _ = myokit.load('example')
    c = myokit_b200.SimulationCUDA(m, p, ncells=16, precision=DP)
    c.set_connections([(i, i + 1, 5.0) for i in range(15)])
    c.set_kernel_options(stage=True)
    assert not (c.kernel_source().kernel_flags & 8)


@pytest.mark.parametrize('lean', [False, True], ids=['slab', 'slab_lean'])
def test_staged_states_row_slabs_equal_the_whole_grid(lean):
    def make(comm):
        kw = {} if comm is None else dict(comm=comm)
        return workloads.c3_hetero(myokit_b200.SimulationCUDA, nx=12, ny=11, **kw)
    opts = dict(EXACT, block=(8, 4))
    whole = make(None)
    whole.set_kernel_options(**opts)
    one = cuda_shim.run_on_host(whole, 3.0, log_interval=0.5)
    for reverse in (False, True):
        out = cuda_shim.run_slabs_on_host(
            make, 3, 3.0, 0.5, dict(opts, slab_lean=lean, stage=True, overlap=True),
            reverse=reverse)
        assert out['halo_error'] == 0
        assert np.array_equal(out['V'], one['V'])
        assert np.array_equal(out['state'], one['state'])


@pytest.mark.parametrize('extra', [
    dict(stage_store=False),
    dict(prefetch_next=0.5, stage_group=(3, 6)),
    dict(stage_store=False, prefetch_next=0.3, select='cheap'),
], ids=['direct_stores', 'next_wave_prefetch', 'both'])
def test_staged_states_variants_bit_for_bit(extra):
    # results written straight to HBM instead of through the staged tile; L2
    # hints for the next wave's tiles (no effect on results)
    def make(cls):
        return workloads.c3_hetero(cls, nx=12, ny=9)
    opts = dict(EXACT, block=(8, 4), stage=True, load_ahead=2, **extra)
    a = make(myokit_b200.SimulationCUDA)
    a.set_kernel_options(**opts)
    src = a.kernel_source()
    assert src.kernel_flags & 8
    if 'prefetch_next' in extra:
        assert 'MKB_TMA_PREFETCH_3D(g.tmap_state' in src.code
    if extra.get('stage_store') is False:
        assert 'MKB_TMA_STORE_3D(g.tmap_state' not in src.code
    got, want, wstate = both(make, opts, 3.0, 0.5, 12, 9)
    assert np.array_equal(got['V'], want['membrane.V'])
    assert np.array_equal(got['state'].ravel(), wstate)


@pytest.mark.parametrize('nx,ny,block,early', [(24, 16, (8, 4), 4), (20, 9, (8, 2), 3), (16, 30, (16, 2), 1)])
def test_staged_tile_loop_kernel_bit_for_bit(nx, ny, block, early):
    # option tile_loop: a fixed number of thread blocks (three here) walks the
    # tiles, the next tile's membrane potentials and first state planes
    # requested while this one is computed; 4-10 tiles per block, ragged rims
    def make(cls):
        return workloads.c3_hetero(cls, nx=nx, ny=ny)
    opts = dict(EXACT, block=block, stage=True, load_ahead=2, tile_loop=True,
                stage_early=early, overlap=False)
    a = make(myokit_b200.SimulationCUDA)
    a.set_kernel_options(**opts)
    src = a.kernel_source()
    assert src.kernel_flags & 8 and src.kernel_flags & 16
    assert 'MKB_CP_ASYNC(&tl_' in src.code
    got, want, wstate = both(make, opts, 3.0, 0.5, nx, ny)
    assert want['membrane.V'].max() > 0
    assert np.array_equal(got['V'], want['membrane.V'])
    assert np.array_equal(got['idiff'], want['membrane.i_diff'])
    assert np.array_equal(got['state'].ravel(), wstate)
    b = make(myokit_b200.SimulationCUDA)
    b.set_kernel_options(**opts)
    rev = cuda_shim.run_on_host(b, 3.0, log_interval=0.5, reverse=True)
    assert np.array_equal(rev['state'], got['state'])


def test_staged_tile_loop_uncoupled_and_homogeneous():
    m, p, _ = myokit.load('example')

    def pop(cls):
        s = cls(m, p, ncells=100, diffusion=False, precision=DP, rl=True)
        s.set_field('ina.gNa', np.linspace(8, 16, 100))
        return s
    opts = dict(EXACT, block=(16, 1), stage=True, tile_loop=True, overlap=False)
    a = pop(myokit_b200.SimulationCUDA)
    a.set_kernel_options(**opts)
    assert a.kernel_source().kernel_flags & 16
    got = cuda_shim.run_on_host(a, 2.0, log_interval=0.5)
    log, ostate = pop(OracleSimulation).run(2.0, log=['engine.time'], log_interval=0.5)
    assert np.array_equal(got['state'].ravel(), np.asarray(ostate))
    # homogeneous conduction, logged intermediary
    got, want, wstate = both(lr91_2d, dict(EXACT, block=(4, 2), stage=True, tile_loop=True,
                                           overlap=False), 5.0, 0.5, 10, 7,
                             inter_log=['ina.INa'])
    assert np.array_equal(got['V'], want['membrane.V'])
    assert np.array_equal(got['idiff'], want['membrane.i_diff'])
    assert np.array_equal(got['inter'][:, 0], want['ina.INa'])
    assert np.array_equal(got['state'].ravel(), wstate)


@pytest.mark.parametrize('block,nx,ny', [((8, 4), 12, 9), ((32, 2), 40, 5), ((64, 1), 70, 3), ((8, 2), 10, 5)])
def test_staged_kernel_neighbours_by_shuffle_bit_for_bit(block, nx, ny):
    # option v_direct: no V tile in shared memory — left / right neighbours from
    # the adjacent lanes of the warp, above / below from memory; blocks narrower
    # and wider than a warp, rows that end inside a warp
    def make(cls):
        return workloads.c3_hetero(cls, nx=nx, ny=ny)
    opts = dict(EXACT, block=block, stage=True, load_ahead=2, v_direct=True)
    a = make(myokit_b200.SimulationCUDA)
    a.set_kernel_options(**opts)
    src = a.kernel_source()
    assert src.kernel_flags & 8 and 'MKB_SHFL_UP(vc, 1)' in src.code
    assert '__shared__ Real tile[' not in src.code
    got, want, wstate = both(make, opts, 3.0, 0.5, nx, ny)
    assert want['membrane.V'].max() > 0
    assert np.array_equal(got['V'], want['membrane.V'])
    assert np.array_equal(got['idiff'], want['membrane.i_diff'])
    assert np.array_equal(got['state'].ravel(), wstate)
    b = make(myokit_b200.SimulationCUDA)
    b.set_kernel_options(**opts)
    rev = cuda_shim.run_on_host(b, 3.0, log_interval=0.5, reverse=True)
    assert np.array_equal(rev['state'], got['state'])


def test_staged_kernel_neighbours_by_shuffle_row_slabs():
    def make(comm):
        kw = {} if comm is None else dict(comm=comm)
        return workloads.c3_hetero(myokit_b200.SimulationCUDA, nx=12, ny=11, **kw)
    opts = dict(EXACT, block=(8, 4))
    whole = make(None)
    whole.set_kernel_options(**opts)
    one = cuda_shim.run_on_host(whole, 3.0, log_interval=0.5)
    for lean in (False, True):
        out = cuda_shim.run_slabs_on_host(
            make, 3, 3.0, 0.5, dict(opts, slab_lean=lean, stage=True, v_direct=True, overlap=True),
            reverse=lean)
        assert out['halo_error'] == 0
        assert np.array_equal(out['V'], one['V'])
        assert np.array_equal(out['state'], one['state'])
