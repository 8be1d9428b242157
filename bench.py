#!/usr/bin/env python
"""
bench.py — cell-steps/s of the SimulationOpenCL hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one time step (fused diffusion stencil + ionic-model update) over
the whole grid of the workload. The workload is BASELINE.json's configs[2],
the one the metric is quoted on: ORd-class 2-D 2048x2048, fp64, Rush-Larsen,
heterogeneous conduction (set_conductance_field) and a per-cell set_field, on
one B200. O'Hara-Rudy CiPA itself is not shipped with the reference; the
reference's own decker-2009.mmt (48 states) stands in and the JSON says so.

The tissue is advanced (untimed) until the stimulus has fired and a wave is
travelling before anything is timed; `config.wave` says how far it got.

Prints ONE JSON line (see the keys below). `value` is measured with all inputs
resident in HBM (CUDA events on the launching stream, inside the library).
`e2e` is the same metric through the public API with HOST buffers, the way the
reference's `run` moves data (state up in sim_init, state down at the end,
openclsim.c:512-562,1185-1190): timed region = `set_state(host array)` +
`run_fields(K dt, ['membrane.V'])` + `state_array()`: the whole state goes up
from page-locked host memory, the logged V field and the whole final state come
back. `e2e.resident` is the second figure: the same call when the state is
left in HBM between runs (this library's normal mode), and `e2e.cold` the
first call of a new simulation starting from the model's initial state.
`--impl reference` times the reference's own CPU arithmetic for the same path
(the reference-rendered kernel compiled as C, oracle/_ref, with OpenMP over all
host cores) on a bounded crop of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "cell-steps/sec (O'Hara-Rudy-class 2D, fp64)"
UNIT = 'cell-steps/s'


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,'
             'clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, device=0):
        self.device = device
        self.proc = None
        self.lines = []
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.device),
                 '--query-gpu=' + self.QUERY, '--format=csv,noheader,nounits',
                 '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return

        def reader():
            for line in self.proc.stdout:
                self.lines.append(line.strip())
        self.thread = threading.Thread(target=reader, daemon=True)
        self.thread.start()

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [],
                    'samples': 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                 'sw_power_cap']
        for line in self.lines:
            parts = [x.strip() for x in line.split(',')]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(names, parts[4:8]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {
            'sm_mhz': sm[len(sm) // 2] if sm else None,
            'sm_max_mhz': max(smax) if smax else None,
            'power_w_max': max(power) if power else None,
            'reasons': sorted(reasons),
            'samples': len(sm),
        }


def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as f:
            return json.load(f), 'measured'
    except (OSError, ValueError):
        return {'hbm_gbs': 6650.0}, 'fallback'


def cpu_baseline(grid, kernel, budget_s=15.0):
    """
    Times the CPU oracle (all host cores, OpenMP) on a bounded crop of the
    workload. kernel: 'ref' (reference-rendered kernel, oracle/_ref) or 'port'.
    Returns (cell_steps_per_s, cores, sample_description, kind).
    """
    from myokit_b200 import workloads
    from oracle.oracle import OracleSimulation
    cores = os.cpu_count() or 1
    kind = 'reference' if kernel == 'ref' else 'port'

    def make():
        return workloads.c3_hetero(
            OracleSimulation, nx=grid, kernel=kernel, openmp=True,
            contract=True, opt='-O3')
    try:
        s = make()
        s.run(2 * 0.005, log=['engine.time'], nthreads=cores)   # build + warm
    except Exception:
        if kernel != 'ref':
            raise
        kernel, kind = 'port', 'port'
        s = make()
        s.run(2 * 0.005, log=['engine.time'], nthreads=cores)
    # Calibrate
    t0 = time.perf_counter()
    s.run(4 * 0.005, log=['engine.time'], nthreads=cores)
    per_step = (time.perf_counter() - t0) / 4
    steps = int(max(8, min(2000, budget_s / max(per_step, 1e-6))))
    s = make()
    s.run(steps * 0.005, log=['engine.time'], nthreads=cores)
    # the native time-step loop only: this repo's Python wrapper around the
    # oracle, and its set-up copies, are not the reference's per-step path
    dt = s.last_loop_seconds
    n_steps = s.last_steps
    value = grid * grid * n_steps / dt
    sample = ('%dx%d crop of the workload (same seeds), %d time steps, %.1f s'
              % (grid, grid, n_steps, dt))
    return value, cores, sample, kind


def run_reference(args, rank, world):
    if rank != 0:
        return
    grid = args.cpu_grid
    t_all = time.perf_counter()
    from myokit_b200 import workloads
    from oracle.oracle import OracleSimulation
    cores = os.cpu_count() or 1
    kernel = 'ref'
    try:
        s = workloads.c3_hetero(OracleSimulation, nx=grid, kernel='ref',
                                openmp=True, contract=True, opt='-O3')
        s.run(0.005, log=['engine.time'], nthreads=cores)
    except Exception:
        kernel = 'port'
        s = workloads.c3_hetero(OracleSimulation, nx=grid, kernel='port',
                                openmp=True, contract=True, opt='-O3')
        s.run(0.005, log=['engine.time'], nthreads=cores)
    # Each "step" = one time step over the crop; W warm-up, K timed. The
    # warm-up is one multi-step call at least as long as the timed one, so that
    # the OpenMP team is up and the working set is paged in (single-step calls
    # left the timed call 1.6x slower than the long cpu_baseline leg).
    s.run(max(args.warmup, args.steps, 50) * 0.005, log=['engine.time'],
          nthreads=cores)
    s.run(args.steps * 0.005, log=['engine.time'], nthreads=cores)
    dt = s.last_loop_seconds     # the native time-step loop (host loop + kernels)
    n_steps = s.last_steps
    value = grid * grid * n_steps / dt
    sample = ('%dx%d crop of the 2048x2048 workload (same seeds), %d time '
              'steps' % (grid, grid, n_steps))
    kind = 'reference' if kernel == 'ref' else 'port'
    out = {
        'impl': 'reference',
        'metric': METRIC, 'value': value, 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': dt / max(n_steps, 1) * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args, cpu=True),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores,
                         'kind': kind, 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'note': ('CPU arithmetic of the reference path: the reference\'s own '
                 'rendered openclsim.cl compiled as C (oracle/_ref) under a '
                 'restated host loop, OpenMP over %d host threads. '
                 'SimulationOpenCL itself cannot run: no OpenCL runtime is '
                 'installable offline. Wall %.1f s.'
                 % (cores, time.perf_counter() - t_all)),
    }
    print(json.dumps(out))


def kernel_profile(key):
    """
    Executed-instruction counts and DRAM traffic of the bench kernel, from the
    committed ncu capture (profiles/c3_kernel_profile.json, written by
    scripts/summarize_profiles.py from an `ncu --set full` report). `key` is
    the hash of the kernel source this run generated: if it differs from the
    profiled kernel's the figures are marked stale instead of being trusted.
    """
    path = os.path.join(ROOT, 'profiles', 'c3_kernel_profile.json')
    try:
        with open(path) as f:
            prof = json.load(f)
    except (OSError, ValueError):
        return None
    prof['stale'] = prof.get('kernel_key') != key
    prof['source'] = 'profiles/c3_kernel_profile.json'
    return prof


def fp64_pipe(prof, peak_ginstr_s, cells_per_launch, kernel_ms):
    if not prof or not prof.get('fp64_instr_per_cell_step'):
        return None
    n = prof['fp64_instr_per_cell_step']
    achieved = n * cells_per_launch / (kernel_ms * 1e-3) / 1e9
    return {'achieved_ginstr_s': achieved, 'peak_ginstr_s': peak_ginstr_s,
            'frac': achieved / peak_ginstr_s if peak_ginstr_s else None,
            'instr_per_cell_step': n,
            'peak_source': 'mkb_measure_peaks, this run (dependent-free DFMA loop)',
            'count_source': prof['source'], 'count_stale': prof['stale']}


def workload_config(args, cpu=False):
    n = args.cpu_grid if cpu else args.grid
    return {
        'workload': ('BASELINE configs[2]: ORd-class 2D %dx%d fp64 '
                     'Rush-Larsen, set_conductance_field + set_field(ikr.Gbar) '
                     'heterogeneity, paced left edge' % (n, n)),
        'model': ('decker-2009.mmt (48 states) — stated proxy: O\'Hara-Rudy '
                  'CiPA is not shipped with the reference'),
        'cells': n * n, 'dt_ms': 0.005, 'scheme': 'rush-larsen',
        'l2_policy': ('working set %.2f GB per step >> 126 MB L2; no flush '
                      'needed' % (n * n * 100 * 8 / 1e9)),
    }


def timed_steps(s, args, world, dist, torch, sampler=None, advance=0):
    """
    (advance +) W untimed steps, then K timed steps, state resident; max over
    ranks of the device time.
    """
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if sampler:
        sampler.start()
    info = s.benchmark_steps(args.steps, warmup=args.warmup + advance)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if sampler else None
    ms = info['device_ms']
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return info, ms, clocks


def max_over_ranks(x, world, dist, torch):
    if world > 1:
        t = torch.tensor([x], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return x


def sharded_check(args, rank, world, local, comm, dist, torch):
    """
    N > 1: the row slabs must give the bits one GPU gives. Every rank runs its
    slab of a 512-row crop through the stimulus; rank 0 also runs the whole
    crop on its GPU alone and compares the V field of the last logged row.
    """
    import numpy as np
    import myokit_b200
    from myokit_b200 import workloads
    nx, ny, steps = 512, 64 * world, 400
    sh = workloads.c3_hetero(myokit_b200.SimulationCUDA, nx=nx, ny=ny,
                             device=local, comm=comm)
    tt, f = sh.run_fields(steps * 0.005, ['membrane.V'], log_interval=(steps - 1) * 0.005)
    mine = torch.from_numpy(np.ascontiguousarray(f['membrane.V'][-1])).cuda()
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    sh.close()
    if rank != 0:
        return None
    whole = workloads.c3_hetero(myokit_b200.SimulationCUDA, nx=nx, ny=ny,
                                device=local)
    tt1, f1 = whole.run_fields(steps * 0.005, ['membrane.V'],
                               log_interval=(steps - 1) * 0.005)
    got = torch.cat(parts).cpu().numpy()
    want = f1['membrane.V'][-1]
    whole.close()
    return {'grid': '%dx%d' % (nx, ny), 'steps': steps,
            'v_range_mV': [float(want.min()), float(want.max())],
            'max_abs_diff_mV': float(np.max(np.abs(got - want))),
            'bit_identical': bool(np.array_equal(got, want))}


def run_ours(args, rank, world):
    import numpy as np
    import torch
    import myokit_b200
    from myokit_b200 import workloads, capi
    local = env_int('LOCAL_RANK', 0)
    dist = None
    comm = None
    if world > 1:
        import torch.distributed as dist
        from myokit_b200 import multigpu
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        comm = multigpu.TorchComm()
    if capi.device_count() < 1:
        raise SystemExit('bench.py: no CUDA device; the product has no CPU path')
    torch.cuda.set_device(local)
    torch.zeros(1, device='cuda')       # the CUDA context exists before any timing

    n = args.grid
    dt = 0.005
    s = workloads.c3_hetero(myokit_b200.SimulationCUDA, nx=n, device=local,
                            comm=comm)
    src = s.kernel_source()
    from myokit_b200.simulation import _cubin_for
    _cubin_for(src)                     # compiled (or found in the cache) before any timing
    n_state = src.n_state
    i_vm = src.i_vm
    alg_bytes = workloads.algorithmic_bytes(n_state, 1, 2, 8)
    cells_total = n * n          # one grid, row slabs over the ranks

    # ---- untimed: advance until the wave travels ----------------------
    # (first call of a new simulation, from the model's initial state: the
    # e2e.cold figure — uniform state, so nothing but the fields goes up)
    t0 = time.perf_counter()
    tt, fields = s.run_fields(args.advance * dt, ['membrane.V'],
                              log_interval=1.0)
    cold_s = max_over_ranks(time.perf_counter() - t0, world, dist, torch)
    cold_info = s.last_run_info()
    x = s.state_array(copy=False)       # page-locked; this rank's cells
    # A wave train: after the advance only the columns next to the paced
    # edge have fired (the front moves ~7 cells per ms). The first
    # `args.wave_period` columns — resting tissue ahead of the front, the
    # upstroke, the plateau behind it — are repeated across the grid, so the
    # timed steps see cells in every phase of the action potential (the
    # libdevice-free kernel has few data-dependent paths left, but the
    # model's own piecewise branches are among them).
    ny_local = x.size // (n_state * n)
    grid_view = x.reshape(ny_local, n, n_state)
    wp = args.wave_period
    if wp and wp < n:
        for c0 in range(wp, n, wp):
            w = min(wp, n - c0)
            grid_view[:, c0:c0 + w, :] = grid_view[:, :w, :]
    v = grid_view[:, :, i_vm]
    wave = np.array([float((v > -60.0).sum()), float(v.size)])
    if world > 1:
        t = torch.from_numpy(wave).cuda()
        dist.all_reduce(t)
        wave = t.cpu().numpy()
    wave = {'advance_steps': int(cold_info['steps']),
            't_ms': float(s.time()),
            'wave_train_period_columns': wp,
            'cells_above_-60mV': wave[0] / wave[1],
            'v_min_max_mV': [float(v.min()), float(v.max())]}
    s.set_state(x)                      # (our own array: adopted, uploaded by the next run)

    # ---- device-resident timing ---------------------------------------
    sampler = ClockSampler(local)
    info, ms, clocks = timed_steps(s, args, world, dist, torch, sampler)
    steps = info['steps']
    value = cells_total * steps / (ms * 1e-3)
    kernel_ms = ms / steps

    # ---- end to end through the public API ----------------------------
    def e2e_call(state_in_out):
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        if state_in_out:
            s.set_state(x)          # our own page-locked array: adopted, uploaded
        tt, f = s.run_fields(args.steps * dt, ['membrane.V'], log_interval=1.0)
        if state_in_out:
            y = s.state_array(copy=False)
            assert y.size == x.size
        sec = max_over_ranks(time.perf_counter() - t0, world, dist, torch)
        return sec, s.last_run_info(), len(tt)
    t_keep = s.time()
    e2e_call(True)                      # warm: log buffers, graphs
    s.set_time(t_keep)
    io_s, io_info, io_rows = e2e_call(True)
    s.set_time(t_keep)
    res_s, res_info, res_rows = e2e_call(False)
    state_bytes = int(x.nbytes)

    # ---- N > 1: the slabs give the bits of one GPU --------------------
    check = None
    if world > 1:
        check = sharded_check(args, rank, world, local, comm, dist, torch)

    # ---- the north star's scaling grid, same process ------------------
    big = None
    if args.scale_grid and args.scale_grid != n:
        s.close()
        del x
        nb = args.scale_grid
        sb = workloads.c3_hetero(myokit_b200.SimulationCUDA, nx=nb,
                                 device=local, comm=comm)
        # (from the model's initial state, broadcast on the device: no host
        # copy of the 25.8 GB state exists at any point)
        binfo, bms, _ = timed_steps(sb, args, world, dist, torch,
                                    advance=args.scale_advance)
        big = {'grid': '%dx%d' % (nb, nb), 'cells': nb * nb,
               'value': nb * nb * binfo['steps'] / (bms * 1e-3), 'unit': UNIT,
               'ms_per_step': bms / binfo['steps'], 'steps': binfo['steps'],
               'advance_steps': args.scale_advance,
               'note': ('BASELINE configs[3] size on the configs[2] model '
                        '(the north star\'s 8192^2 ORd-class strong-scaling '
                        'target); device-resident, max over ranks')}
        sb.close()

    if rank != 0:
        return
    peaks, peaks_src = load_peaks()
    # per GPU: each launch covers this rank's slab
    achieved = alg_bytes * (n * n / world) / (kernel_ms * 1e-3) / 1e9
    peak = float(peaks.get('hbm_gbs', 6650.0))
    prof = kernel_profile(src.key()) if world == 1 else None
    try:
        pipe_peak = capi.measure_peaks(local)['fp64_fma_ginstr_s']
    except Exception:
        pipe_peak = None

    cpu = None
    if world == 1 and not args.no_cpu:
        try:
            v_, cores, sample, kind = cpu_baseline(args.cpu_grid, 'ref')
            cpu = {'value': v_, 'unit': UNIT, 'cores': cores, 'kind': kind,
                   'sample': sample}
        except Exception as e:     # keep the GPU line even if gcc is missing
            cpu = {'value': None, 'unit': UNIT, 'cores': 0, 'kind': 'port',
                   'sample': 'failed: %s' % e}

    def per_step(info, key):
        return info[key] / max(info['steps'], 1)

    out = {
        'metric': METRIC, 'value': value, 'unit': UNIT,
        'n_gpus': world, 'steps': steps, 'warmup': args.warmup,
        'ms_per_step': kernel_ms,
        'higher_is_better': True,
        'scaling': 'strong',
        'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': dict(workload_config(args), wave=wave, parallelism=(
            'single GPU' if world == 1 else
            '%d row slabs of %d rows, one process per GPU; ghost rows of V '
            'pushed by the step kernel into the neighbour over NVLink '
            '(CUDA IPC peer stores + arrival flags), no collective on the '
            'step path' % (world, n // world))),
        'clocks': clocks,
        'e2e': {'value': cells_total * io_info['steps'] / io_s, 'unit': UNIT,
                'h2d_bytes_per_step': per_step(io_info, 'h2d_bytes'),
                'd2h_bytes_per_step': (io_info['d2h_bytes'] + state_bytes)
                / max(io_info['steps'], 1),
                'seconds': io_s, 'steps': io_info['steps'],
                'api': 'SimulationCUDA.set_state + run_fields + state_array',
                'log_rows': io_rows,
                'state_bytes_each_way': state_bytes,
                'host_seconds': io_info.get('host_seconds'),
                'note': ('every call moves the whole state up from page-'
                         'locked host memory and back down, as the reference '
                         'does in sim_init / at the end of sim_step; with K '
                         'short this is mostly PCIe time'),
                # the same call with the state left in HBM between runs
                'resident': {
                    'value': cells_total * res_info['steps'] / res_s,
                    'unit': UNIT, 'seconds': res_s,
                    'h2d_bytes_per_step': per_step(res_info, 'h2d_bytes'),
                    'd2h_bytes_per_step': per_step(res_info, 'd2h_bytes'),
                    'log_rows': res_rows,
                    'host_seconds': res_info.get('host_seconds')},
                # the first call on a new simulation (initial state broadcast
                # on the device; fields, kernel load, allocation included)
                'cold': {'value': cells_total * cold_info['steps'] / cold_s,
                         'unit': UNIT, 'seconds': cold_s,
                         'steps': cold_info['steps'],
                         'h2d_bytes': cold_info['h2d_bytes'],
                         'd2h_bytes': cold_info['d2h_bytes'],
                         'host_seconds': cold_info.get('host_seconds')}},
        'gpu_launches': info['kernel_launches'],
        'roofline': {
            'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
            'frac': achieved / peak,
            # dram__bytes_read + dram__bytes_write of one launch, from the
            # committed ncu --set full capture of this kernel
            'traffic': (prof['dram_bytes_per_launch']
                        if prof and not prof['stale'] and n == 2048 else None),
            'traffic_source': prof['source'] if prof else None,
            'peak_source': peaks_src + ' (MEASURED_PEAKS.json hbm_gbs)',
            'kernel': 'mkb_cell_step',
            'algorithmic_bytes_per_cell_step': alg_bytes,
            'note': ('fused stencil + cell update: the HBM floor (800 B per '
                     'cell-step) and the FP64-pipe floor (fp64_pipe below) of '
                     'this kernel lie within 5 % of each other (DESIGN.md §4.1)'),
            # The binding ceiling: FP64-pipe instructions executed per
            # cell-step (committed ncu capture) against the FP64 FMA issue
            # rate measured in this run.
            'fp64_pipe': fp64_pipe(prof, pipe_peak, n * n / world, kernel_ms),
        },
        'cpu_baseline': cpu,
    }
    if check is not None:
        out['sharded_check'] = check
    if big is not None:
        out['scaling_8192'] = big
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--grid', type=int, default=2048)
    ap.add_argument('--cpu-grid', type=int, default=256)
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--wave-period', type=int, default=128,
                    help='columns of the advanced tissue repeated across the grid (0: off)')
    ap.add_argument('--advance', type=int, default=4000,
                    help='untimed time steps before anything is timed')
    ap.add_argument('--scale-grid', type=int, default=8192,
                    help='second grid timed in the same process (0: none)')
    ap.add_argument('--scale-advance', type=int, default=300)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    rank = env_int('RANK', 0)
    world = env_int('WORLD_SIZE', 1)
    if args.impl == 'reference':
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world)


if __name__ == '__main__':
    main()
