"""
Host-side logic of the multi-GPU (row-slab) path on CPU: world_size 2 over
gloo, one process per rank, exactly as torchrun would start them. Checks the
communicator, the slab partition, slicing of global inputs, the local log
table, and that running without a GPU still fails loudly on every rank.
"""
import os
import subprocess
import sys
import textwrap

from myokit_b200 import multigpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import os, sys, json
    sys.path.insert(0, %r)
    import numpy as np
    import torch.distributed as dist
    dist.init_process_group('gloo')
    import myokit_b200, myokit
    from myokit_b200 import multigpu, capi
    comm = multigpu.TorchComm()
    assert comm.size == 2
    got = comm.allgather(('rank', comm.rank, b'x' * 64))
    assert [g[1] for g in got] == [0, 1]
    comm.barrier()
    # NaN-halt agreement: one small all-reduce over the host group
    assert comm.any(False) is False
    assert comm.any(comm.rank == 1) is True
    # set_state on this rank's own array is adopted, a uniform state stays one cell's values

    m, p, _ = myokit.load('example')
    nx, ny = 6, 5
    s = myokit_b200.SimulationCUDA(m, p, ncells=(nx, ny), precision=myokit.DOUBLE_PRECISION, comm=comm)
    x0, snx, y0, sny = s.local_shape()
    rows = multigpu.slab_rows(ny, 2)
    assert (y0, y0 + sny) == rows[comm.rank] and (x0, snx) == (0, nx)
    n = m.count_states()
    assert len(s.state_array()) == sny * nx * n

    # global state in, local slab kept
    full = np.arange(nx * ny * n, dtype=float)
    s.set_state(full)
    want = full.reshape(ny, nx, n)[y0:y0 + sny].reshape(-1)
    assert np.array_equal(s.state_array(), want)
    mine = s.state_array(copy=False)
    s.set_state(mine)                       # our own array handed back: no copy
    assert s.state_array(copy=False) is mine
    u1 = myokit_b200.SimulationCUDA(m, p, ncells=(nx, ny), precision=myokit.DOUBLE_PRECISION, comm=comm)
    assert u1._state_full is None and len(u1._state_cell) == n      # uniform: never tiled
    u1.set_state(list(range(n)))
    assert u1._state_full is None and u1.state(2, y0) == list(map(float, range(n)))
    # single-cell access is global-indexed; remote cells raise
    own = (2, y0)
    other = (2, rows[1 - comm.rank][0])
    assert s.state(*own) == list(full.reshape(ny, nx, n)[own[1], own[0]])
    try:
        s.state(*other)
        raise SystemExit('expected IndexError')
    except IndexError:
        pass
    # fields are given globally and sliced per rank
    g = np.arange(nx * ny, dtype=float).reshape(ny, nx)
    s.set_field('ina.gNa', g)
    assert np.array_equal(s._local_slice(s._fields[s._model.get('ina.gNa')]), g[y0:y0 + sny].reshape(-1))
    # the log table only holds this rank's cells, numbered locally
    log = myokit.prepare_log(['engine.time', 'membrane.V'], s._model, dims=(nx, ny), global_vars=['engine.time', 'engine.pace'])
    keys, kinds, index = s._log_table(log, [])
    assert keys[0] == 'engine.time' and len(keys) == 1 + sny * nx
    assert '0.%%d.membrane.V' %% y0 in keys and ('0.%%d.membrane.V' %% other[1]) not in keys
    k = keys.index('3.%%d.membrane.V' %% (y0 + 1))
    assert kinds[k] == capi.LOG_STATE and index[k] == (3 + 1 * nx) * n + 0
    # slab kernels are generated for coupled grids only
    assert 'mkb_wait_flag(g.flag_lo' in s.kernel_source().code
    u = myokit_b200.SimulationCUDA(m, p, ncells=10, diffusion=False, comm=comm)
    assert u.local_shape() == ((0, 5, 0, 1), (5, 5, 0, 1))[comm.rank]
    assert 'g.flag_lo' not in u.kernel_source().code
    # a coupled 1-d cable cannot be cut; a set_connections graph can
    c = myokit_b200.SimulationCUDA(m, p, ncells=10, comm=comm, precision=myokit.DOUBLE_PRECISION)
    try:
        c.kernel_source()
        raise SystemExit('expected ValueError')
    except ValueError as e:
        assert 'cannot be sharded' in str(e)
    edges = [(i, i + 1, 2.0 + i) for i in range(9)] + [(0, 7, 0.5), (2, 9, 0.25), (6, 1, 0.75)]
    c.set_connections(edges)
    assert 'g.ghost' in c.kernel_source().code
    li, lj, lg, ghosts = c._partition_graph()
    x0 = 5 * comm.rank
    # every global edge touching this block appears once, local endpoint first
    want = [(a, b, g) for (a, b, g) in [(min(a, b), max(a, b), g) for a, b, g in edges]
            if x0 <= a < x0 + 5 or x0 <= b < x0 + 5]
    assert len(li) == len(want)
    expect_ghosts = sorted(set(v for a, b, g in want for v in (a, b) if not x0 <= v < x0 + 5))
    assert list(ghosts) == expect_ghosts
    for k, (a, b, gg) in enumerate(want):
        loc, rem = (a, b) if x0 <= a < x0 + 5 else (b, a)
        assert li[k] == loc - x0 and lg[k] == gg
        if x0 <= rem < x0 + 5:
            assert lj[k] == rem - x0
        else:
            assert lj[k] == 5 + expect_ghosts.index(rem)

    # gather_rows reassembles slabs
    loc = g[y0:y0 + sny][None]
    assert np.array_equal(multigpu.gather_rows(comm, loc), g[None])

    # no GPU here: the run must fail on every rank, not fall back
    if capi.device_count() == 0:
        try:
            s.run(1)
            raise SystemExit('expected BackendError')
        except capi.BackendError as e:
            assert 'no CPU fallback' in str(e)
    dist.barrier()
    sys.stdout.write('RANK_OK_%%d%%s' %% (comm.rank, os.linesep))   # one write: ranks share the pipe
    sys.stdout.flush()
''') % ROOT


def test_two_ranks_over_gloo(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(WORKER)
    env = dict(os.environ)
    env.pop('CUDA_VISIBLE_DEVICES', None)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
           '--nproc-per-node=2', '--master-addr', '127.0.0.1',
           '--master-port', '29531', str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert 'RANK_OK_0' in r.stdout and 'RANK_OK_1' in r.stdout


def test_thread_comm_and_slab_rows():
    assert multigpu.slab_rows(10, 4) == [(0, 2), (2, 5), (5, 7), (7, 10)]

    def work(comm):
        vals = comm.allgather(comm.rank * 10)
        comm.barrier()
        return vals
    out = multigpu.run_threads(3, work)
    assert out == [[0, 10, 20]] * 3

    def boom(comm):
        if comm.rank == 1:
            raise KeyError('x')
        comm.barrier()
    try:
        multigpu.run_threads(2, boom)
        assert False
    except KeyError:
        pass


def test_slab_kernel_variants_compile():
    """Both forms of the row-slab kernel build for sm_100a (NVRTC, no GPU)."""
    import myokit
    import myokit_b200
    from myokit_b200 import capi

    m, p, _ = myokit.load('example')
    out = {}

    def work(comm):
        for lean in (False, True):
            s = myokit_b200.SimulationCUDA(
                m, p, ncells=(70, 9), precision=myokit.DOUBLE_PRECISION,
                comm=comm)
            s.set_kernel_options(slab_lean=lean)
            if comm.rank == 0:
                out[lean] = s.kernel_source()
    multigpu.run_threads(2, work)
    plain, lean = out[False].code, out[True].code
    # default form: inactive threads skip the cell model inside a region
    assert '    if (active) {\n' in plain and 'mkb_slab_pos<MKB_BY>(' not in plain
    # lean form: they leave, and the publication re-reads its coordinates
    assert '    if (!active) return;\n' in lean
    assert 'const MkbSlabPos q = mkb_slab_pos<MKB_BY>((unsigned int)g.ny);' in lean
    for src in (out[False], out[True]):
        assert 'peer_lo_flag_hi' in src.code
        cubin, log = capi.jit_compile(src.code, src.options)
        assert len(cubin) > 10000
