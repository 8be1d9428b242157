"""Throughput of the other BASELINE configurations (device-resident timing)."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np
import myokit_b200, myokit
from myokit_b200 import workloads

_S = myokit_b200.SimulationCUDA
WORLD = int(os.environ.get('WORLD_SIZE', '1'))
RANK = int(os.environ.get('RANK', '0'))
COMM = None
if WORLD > 1:
    # one process per GPU under torchrun: shard every workload
    import torch
    import torch.distributed as dist
    from myokit_b200 import multigpu
    LOCAL = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(LOCAL)
    dist.init_process_group('nccl', device_id=torch.device('cuda', LOCAL))
    COMM = multigpu.TorchComm()


def S(*args, **kw):
    if COMM is not None:
        kw.setdefault('device', LOCAL)
        kw.setdefault('comm', COMM)
    return _S(*args, **kw)


def report(name, s, steps, warmup=20, **opts):
    if opts:
        s.set_kernel_options(**opts)
    info = s.benchmark_steps(steps, warmup=warmup)
    ms = info['device_ms'] / info['steps']
    cells = info['cells']
    if COMM is not None:
        # whole job: all cells / slowest rank
        got = COMM.allgather((ms, cells))
        ms = max(g[0] for g in got)
        cells = sum(g[1] for g in got)
        if RANK != 0:
            return
        name = '[%d GPUs] %s' % (WORLD, name)
    print('%-44s %9d cells  %8.4f ms/step  %.3e cell-steps/s  (%d launches/rank)' % (
        name, cells, ms, cells / ms * 1e3, info['kernel_launches']), flush=True)


which = sys.argv[1:] or ['c1', 'c2', 'c4', 'c5']
if 'c1' in which:
    report('C1 LR91 1-D 128 fp64', workloads.c1_cable(S, 128), 20000)
    report('C1-like LR91 1-D 16384 fp64', workloads.c1_cable(S, 16384), 5000)
    # all unlogged steps of a schedule chunk in one launch (one thread block)
    report('C1 LR91 1-D 128 fp64 persistent', workloads.c1_cable(S, 128), 20000, persistent=True)
    report('C1 128 persistent + select', workloads.c1_cable(S, 128), 20000, persistent=True, select=True)
    report('C1 128 persistent + select + estrin', workloads.c1_cable(S, 128), 20000, persistent=True, select=True, fast_exp='estrin')
    report('C1 128 persistent + select + estrin + div_parallel', workloads.c1_cable(S, 128), 20000, persistent=True, select=True, fast_exp='estrin', div_cubic=False, div_parallel=True)
    report('C1-like LR91 1-D 1024 fp64 persistent + select', workloads.c1_cable(S, 1024), 20000, persistent=True, select=True)
    report('C1-like LR91 1-D 1024 fp64', workloads.c1_cable(S, 1024), 20000)
if 'c2' in which:
    report('C2 LR91 512^2 fp32', workloads.c2_planar(S, 512), 5000)
    report('C2 LR91 512^2 fp32 b32x8', workloads.c2_planar(S, 512), 5000, block=(32, 8))
    report('C2 LR91 512^2 fp32 b128x2', workloads.c2_planar(S, 512), 5000, block=(128, 2))
    report('C2 LR91 2048^2 fp32', workloads.c2_planar(S, 2048), 500)
if 'c4' in which:
    report('C4-proxy LR91 8192^2 fp32', workloads.c2_planar(S, 8192), 50, warmup=5)
    # (measured in round 1 and dropped: fast_div=True is the default already; min_blocks=4 is 10 % slower)
    if os.environ.get('MKB_TEST_EXPERIMENTAL'):
        # prepared without GPU time left to measure it: 6-instruction expf
        report('C4-proxy LR91 8192^2 fp32 exp=ex2', workloads.c2_planar(S, 8192), 50, warmup=5, fast_exp='ex2')
if 'stencil' in which:
    for prec, name in ((myokit.SINGLE_PRECISION, 'fp32'), (myokit.DOUBLE_PRECISION, 'fp64')):
        rs = 4 if name == 'fp32' else 8
        variants = [dict(stream=False), dict(stream=True), dict(stream=True, min_blocks=3),
                    dict(stream=True, min_blocks=4), dict(stream=True, block=(32, 4)),
                    dict(stream=True, block=(32, 4), min_blocks=4),
                    dict(stream=True, block=(32, 8), rows_per_thread=2, min_blocks=4),
                    dict(stream=True, block=(32, 16), rows_per_thread=2, min_blocks=2)]
        if name == 'fp64':
            variants += [dict(stream=True, cells_per_thread=4), dict(stream=True, cells_per_thread=4, min_blocks=3)]
        for opts in variants:
            s = workloads.stencil_only(S, 8192, 4096, precision=prec)
            s.set_kernel_options(**opts)
            try:
                info = s.benchmark_steps(50, warmup=5)
            except Exception as e:
                print('stencil-only %s %s failed: %s' % (name, opts, str(e)[:300]), flush=True)
                continue
            ms = info['device_ms'] / info['steps']
            gbs = 2 * rs * info['cells'] / ms / 1e6
            print('stencil-only 8192x4096 %s %-60s %8.4f ms/step  %7.1f GB/s  (%.1f%% of 6392.8)' % (
                name, opts, ms, gbs, 100 * gbs / 6392.8), flush=True)
    s = workloads.stencil_only(S, 8192, precision=myokit.SINGLE_PRECISION, hetero=True)
    info = s.benchmark_steps(50, warmup=5)
    ms = info['device_ms'] / info['steps']
    gbs = 4 * 4 * info['cells'] / ms / 1e6
    print('stencil-only 8192^2 fp32 hetero           %8.4f ms/step  %7.1f GB/s  (%.1f%% of 6392.8)' % (ms, gbs, 100 * gbs / 6392.8))
if 'mesh' in which:
    t0 = time.time()
    s = workloads.c5_mesh(S)
    if RANK == 0:
        print('mesh built in %.1f s: %d cells, %d edges' % (time.time() - t0, s._nx, len(s._connections[0])))
    report('C5-ii LR91 fp32 4.2M-node fibre mesh', s, 100, warmup=5)
if 'c5' in which:
    m = workloads.data_model('decker-2009.mmt')
    p = myokit.pacing.blocktrain(period=1000, duration=2, offset=1)
    n = 1 << 20
    s = S(m, p, ncells=n, diffusion=False, precision=myokit.DOUBLE_PRECISION, rl=True)
    rng = np.random.default_rng(42)
    for var, base in (('ikr.Gbar', 0.0138542), ('ina.Gbar', 9.075), ('ik1.Gbar', 0.5)):
        s.set_field(var, base * (1 - rng.uniform(0, 1, n)))
    report('C5-i decker 1M uncoupled fp64 RL', s, 200, warmup=5)
if 'meshtrace' in which:
    # Where a partitioned-mesh step goes: kernel table from the CUPTI tracer
    # (diagnostic only; numbers under a tracer are never bench values).
    import torch
    from torch.profiler import profile, ProfilerActivity
    nz = 128 if WORLD > 1 else 128 // int(os.environ.get('MESH_SPLIT', '2'))
    s = workloads.c5_mesh(S, nz=nz)
    report('mesh nz=%d (untraced)' % nz, s, 100, warmup=5)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        s.benchmark_steps(50, warmup=0)
        torch.cuda.synchronize()
    if RANK == 0:
        ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
        ev.sort(key=lambda e: e.time_range.start)
        by = {}
        for e in ev:
            by.setdefault(e.name[:60], []).append(e.time_range.end - e.time_range.start)
        for k, v in by.items():
            print('  %-60s n=%4d  mean %8.2f us' % (k, len(v), sum(v) / len(v)))
        if ev:
            span = ev[-1].time_range.end - ev[0].time_range.start
            busy = sum(e.time_range.end - e.time_range.start for e in ev)
            print('  span %.1f us, kernels %.1f us, gaps %.1f us' % (span, busy, span - busy))
