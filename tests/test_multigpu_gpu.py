"""
Row-slab sharding on real devices: N slabs driven by N threads of one process
(ThreadComm, direct peer pointers). With fewer GPUs than slabs the slabs share
a device — the exchange protocol (arrival flags, three ghost-row slots, peer
stores) is the same, so one B200 already exercises it; a multi-GPU box then
adds NVLink under the same code. The sharded result must equal the unsharded
one bit for bit: same kernel arithmetic, same V rows, only fetched differently.
"""
import numpy as np
import pytest

import myokit_b200
from myokit_b200 import capi, multigpu, workloads
import myokit

pytestmark = pytest.mark.gpu
DP = myokit.DOUBLE_PRECISION
SP = myokit.SINGLE_PRECISION


def devices(n):
    have = capi.device_count()
    return [r % max(have, 1) for r in range(n)]


def sharded(make, nslab, duration, variables, log_interval):
    devs = devices(nslab)

    def target(comm):
        s = make(devs[comm.rank], comm)
        t, f = s.run_fields(duration, variables, log_interval=log_interval)
        full = dict((k, multigpu.gather_rows(comm, v)) for k, v in f.items())
        state = comm.allgather(s.state_array())
        return t, full, np.concatenate(state), s.last_run_info()
    return multigpu.run_threads(nslab, target)


@pytest.mark.parametrize('nslab', [2, 3])
def test_slabs_equal_single_gpu_homogeneous_fp64(nslab):
    m, _, _ = myokit.load('example')
    p = myokit.pacing.blocktrain(duration=2, offset=1, period=1000)
    nx, ny = 70, 26

    def make(device, comm):
        s = myokit_b200.SimulationCUDA(m, p, ncells=(nx, ny), precision=DP,
                                       device=device, comm=comm)
        s.set_conductance(9, 6)
        s.set_paced_cells(4, 7, 0, 9)      # straddles a slab boundary
        return s
    ref = make(0, None)
    t0, f0 = ref.run_fields(6, ['membrane.V', 'membrane.i_diff'], 0.5)
    assert f0['membrane.V'][-1].max() > 0
    out = sharded(make, nslab, 6, ['membrane.V', 'membrane.i_diff'], 0.5)
    for t, f, state, info in out:
        assert np.array_equal(t, t0)
        assert np.array_equal(f['membrane.V'], f0['membrane.V'])
        assert np.array_equal(f['membrane.i_diff'], f0['membrane.i_diff'])
        assert np.array_equal(state, ref.state_array())
        assert info['steps'] == ref.last_run_info()['steps']


def test_slabs_equal_single_gpu_c3_hetero_fields():
    nx, ny = 64, 40

    def make(device, comm):
        return workloads.c3_hetero(myokit_b200.SimulationCUDA, nx=nx, ny=ny,
                                   device=device, comm=comm)
    ref = make(0, None)
    t0, f0 = ref.run_fields(5, ['membrane.V'], 0.5)
    out = sharded(make, 4, 5, ['membrane.V'], 0.5)
    for t, f, state, info in out:
        assert np.array_equal(f['membrane.V'], f0['membrane.V'])
        assert np.array_equal(state, ref.state_array())


def test_slabs_fp32_many_steps_and_datalog_keys():
    m, _, _ = myokit.load('example')
    p = myokit.pacing.blocktrain(duration=2, offset=1, period=1000)
    nx, ny = 40, 16

    def make(device, comm):
        s = myokit_b200.SimulationCUDA(m, p, ncells=(nx, ny), precision=SP,
                                       device=device, comm=comm)
        s.set_conductance(8, 8)
        s.set_paced_cells(3, ny, 0, 0)
        return s
    ref = make(0, None)
    d0 = ref.run(12, log=['engine.time', 'membrane.V'], log_interval=1)
    devs = devices(2)

    def target(comm):
        s = make(devs[comm.rank], comm)
        d = s.run(12, log=['engine.time', 'membrane.V'], log_interval=1)
        return dict((k, np.array(v)) for k, v in d.items()), s.local_shape()
    out = multigpu.run_threads(2, target)
    seen = set()
    for d, (x0, snx, y0, sny) in out:
        assert len(d) == 1 + snx * sny       # only this rank's cells
        for k, v in d.items():
            assert np.array_equal(v, np.array(d0[k])), k
            seen.add(k)
    assert seen == set(d0.keys())


def test_uncoupled_population_shards_without_exchange():
    m, _, _ = myokit.load('example')
    p = myokit.pacing.blocktrain(duration=2, offset=1, period=1000)
    n = 101
    g = np.linspace(6, 16, n)

    def make(device, comm):
        s = myokit_b200.SimulationCUDA(m, p, ncells=n, diffusion=False,
                                       precision=DP, rl=True, device=device,
                                       comm=comm)
        s.set_field('ina.gNa', g)
        # (101 cells on one GPU would take the one-block persistent kernel with
        # its Estrin exp: same kernel arithmetic on both sides for a bit-wise test)
        s.set_kernel_options(persistent=False)
        return s
    ref = make(0, None)
    t0, f0 = ref.run_fields(4, ['membrane.V'], 0.5)
    devs = devices(3)

    def target(comm):
        s = make(devs[comm.rank], comm)
        t, f = s.run_fields(4, ['membrane.V'], 0.5)
        return comm.allgather(f['membrane.V'])
    out = multigpu.run_threads(3, target)
    assert np.array_equal(np.concatenate(out[0], axis=-1), f0['membrane.V'])


@pytest.mark.parametrize('nparts,precision', [(2, DP), (3, DP), (4, SP)])
def test_partitioned_connection_graph_equals_single_gpu(nparts, precision):
    # set_connections graph cut into contiguous id blocks: ghost cells are
    # pushed by their owners every step. Same per-cell summation order as on
    # one GPU (edge-list order), so the same bits.
    m, _, _ = myokit.load('example')
    p = myokit.pacing.blocktrain(duration=2, offset=1, period=1000)
    n, edges = workloads.fibre_mesh(12, 5, 4, extra=0.08, seed=3)

    def make(device, comm):
        s = myokit_b200.SimulationCUDA(m, p, ncells=n, precision=precision,
                                       device=device, comm=comm)
        s.set_connections(edges)
        s.set_paced_cell_list([0, 1, 2, 12, 13, 60, 61])
        return s
    ref = make(0, None)
    t0, f0 = ref.run_fields(8, ['membrane.V', 'membrane.i_diff'], 0.5)
    V0 = f0['membrane.V']
    assert V0[-1].max() > 0 and np.any((V0.min(axis=1) < -40) & (V0.max(axis=1) > -20))
    devs = devices(nparts)

    def target(comm):
        s = make(devs[comm.rank], comm)
        t, f = s.run_fields(8, ['membrane.V', 'membrane.i_diff'], 0.5)
        # and a second run on the resident state (re-arm + re-seed)
        t2, f2 = s.run_fields(2, ['membrane.V'], 0.5)
        return (comm.allgather(f['membrane.V']),
                comm.allgather(f['membrane.i_diff']),
                comm.allgather(f2['membrane.V']),
                comm.allgather(s.state_array()))
    out = multigpu.run_threads(nparts, target)
    t2, g2 = ref.run_fields(2, ['membrane.V'], 0.5)
    V, I, V2, S = out[0]
    assert np.array_equal(np.concatenate(V, axis=-1), f0['membrane.V'])
    assert np.array_equal(np.concatenate(I, axis=-1), f0['membrane.i_diff'])
    assert np.array_equal(np.concatenate(V2, axis=-1), g2['membrane.V'])
    assert np.array_equal(np.concatenate(S), ref.state_array())


@pytest.mark.parametrize('nslab', [2, 3])
def test_lean_slab_kernel_equals_single_gpu(nslab):
    # grid sizes that leave threads outside the grid in every block row/column
    m, _, _ = myokit.load('example')
    p = myokit.pacing.blocktrain(duration=2, offset=1, period=1000)
    nx, ny = 70, 26

    def make(device, comm):
        s = myokit_b200.SimulationCUDA(m, p, ncells=(nx, ny), precision=DP,
                                       device=device, comm=comm)
        s.set_conductance(9, 6)
        s.set_paced_cells(4, 7, 0, 9)
        s.set_kernel_options(slab_lean=True)
        return s
    ref = make(0, None)
    t0, f0 = ref.run_fields(6, ['membrane.V'], 0.5)
    out = sharded(make, nslab, 6, ['membrane.V'], 0.5)
    for t, f, state, info in out:
        assert np.array_equal(f['membrane.V'], f0['membrane.V'])
        assert np.array_equal(state, ref.state_array())
