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

Prints ONE JSON line (see the keys below). `value` is measured with all inputs
resident in HBM (CUDA events on the launching stream, inside the library);
`e2e` is the same metric through the public `SimulationCUDA.run_fields` call
with host buffers (state upload, log + final-state download inside the timed
region). `--impl reference` times the reference's own CPU arithmetic for the
same path (the reference-rendered kernel compiled as C, oracle/_ref, with
OpenMP over all host cores) on a bounded crop of the same workload.
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
    # the native call only: this repo's Python wrapper around the oracle is
    # not the reference's overhead
    dt = s.last_run_seconds
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
    # Each "step" = one time step over the crop; W warm-up, K timed
    for _ in range(args.warmup):
        s.run(0.005, log=['engine.time'], nthreads=cores)
    s.run(args.steps * 0.005, log=['engine.time'], nthreads=cores)
    dt = s.last_run_seconds      # the native call (host loop + kernels) only
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


FP64_INSTR_PER_CELL_STEP = 2255      # DFMA + DMUL + DADD + DSETP per cell-step, profiles/r01_opmix_c3.txt
FP64_PEAK_GINSTR_S = 17105.3         # mkb_measure_peaks, r01


def fp64_pipe(cells_per_launch, kernel_ms):
    achieved = FP64_INSTR_PER_CELL_STEP * cells_per_launch / (kernel_ms * 1e-3) / 1e9
    return {'achieved_ginstr_s': achieved, 'peak_ginstr_s': FP64_PEAK_GINSTR_S,
            'frac': achieved / FP64_PEAK_GINSTR_S,
            'instr_per_cell_step': FP64_INSTR_PER_CELL_STEP}


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


def run_ours(args, rank, world):
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

    n = args.grid
    s = workloads.c3_hetero(myokit_b200.SimulationCUDA, nx=n, device=local,
                            comm=comm)
    src = s.kernel_source()
    n_state = src.n_state
    alg_bytes = workloads.algorithmic_bytes(n_state, 1, 2, 8)

    # ---- device-resident timing -------------------------------------
    # A first short call compiles / loads the kernel and touches all memory
    s.benchmark_steps(2, warmup=1)
    sampler = ClockSampler(local)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    info = s.benchmark_steps(args.steps, warmup=args.warmup)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()
    ms = info['device_ms']
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    steps = info['steps']
    cells_total = n * n          # one grid, row slabs over the ranks
    value = cells_total * steps / (ms * 1e-3)
    kernel_ms = ms / steps

    # ---- end to end through the public API ----------------------------
    # First call: uploads the state (1.6 GB of host doubles) and leaves it
    # resident; timed call: the same public call again, continuing the run —
    # host-side pacing schedule in, logged V field out, every step through the
    # library's host loop.
    s2 = workloads.c3_hetero(myokit_b200.SimulationCUDA, nx=n, device=local,
                             comm=comm)
    t0 = time.perf_counter()
    # (long enough that the library has built its CUDA graphs — they are
    # instantiated on the first batch of 64 plain steps —
    # and at least as long as the timed call, so that the library's log
    # buffers already have their final size)
    s2.run_fields(max(args.warmup, 200, args.steps) * 0.005, ['membrane.V'],
                  log_interval=1.0)
    cold_s = time.perf_counter() - t0
    cold_info = s2.last_run_info()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    tt, fields = s2.run_fields(args.steps * 0.005, ['membrane.V'],
                               log_interval=1.0)
    e2e_s = time.perf_counter() - t0
    i2 = s2.last_run_info()
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = cells_total * i2['steps'] / e2e_s

    if rank != 0:
        return
    peaks, peaks_src = load_peaks()
    # per GPU: each launch covers this rank's slab
    achieved = alg_bytes * (n * n / world) / (kernel_ms * 1e-3) / 1e9
    peak = float(peaks.get('hbm_gbs', 6650.0))

    cpu = None
    if world == 1 and not args.no_cpu:
        try:
            v, cores, sample, kind = cpu_baseline(args.cpu_grid, 'ref')
            cpu = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': kind,
                   'sample': sample}
        except Exception as e:     # keep the GPU line even if gcc is missing
            cpu = {'value': None, 'unit': UNIT, 'cores': 0, 'kind': 'port',
                   'sample': 'failed: %s' % e}

    out = {
        'metric': METRIC, 'value': value, 'unit': UNIT,
        'n_gpus': world, 'steps': steps, 'warmup': args.warmup,
        'ms_per_step': kernel_ms,
        'higher_is_better': True,
        'scaling': 'strong',
        'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': dict(workload_config(args), parallelism=(
            'single GPU' if world == 1 else
            '%d row slabs of %d rows, one process per GPU; ghost rows of V '
            'pushed by the step kernel into the neighbour over NVLink '
            '(CUDA IPC peer stores + arrival flags), no collective on the '
            'step path' % (world, n // world))),
        'clocks': clocks,
        'e2e': {'value': e2e_value, 'unit': UNIT,
                'h2d_bytes_per_step': i2['h2d_bytes'] / max(i2['steps'], 1),
                'd2h_bytes_per_step': i2['d2h_bytes'] / max(i2['steps'], 1),
                'seconds': e2e_s, 'api': 'SimulationCUDA.run_fields',
                'log_rows': int(len(tt)),
                'host_seconds': i2.get('host_seconds'),
                # the first call on a new simulation, state upload included
                'cold': {'value': cells_total * cold_info['steps'] / cold_s,
                         'unit': UNIT, 'seconds': cold_s,
                         'steps': cold_info['steps'],
                         'h2d_bytes': cold_info['h2d_bytes'],
                         'd2h_bytes': cold_info['d2h_bytes']},
                'note': ('second run_fields() call on the same simulation: '
                         'the state stays in HBM between runs (as the '
                         'reference keeps it in a Python list), so the timed '
                         'call moves the pacing schedule in and the logged V '
                         'field out; the first call, which also uploads the '
                         '%.2f GB initial state, took %.2f s for %d steps'
                         % (cold_info['h2d_bytes'] / 1e9, cold_s,
                            cold_info['steps']))},
        'gpu_launches': info['kernel_launches'],
        'roofline': {
            'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
            'frac': achieved / peak,
            # dram__bytes_read + dram__bytes_write of one launch, from the
            # ncu --set full capture of this workload (profiles/r01_summary.md)
            'traffic': 3.412e9 if (n == 2048 and world == 1) else None,
            'peak_source': peaks_src + ' (MEASURED_PEAKS.json hbm_gbs)',
            'kernel': 'mkb_cell_step',
            'algorithmic_bytes_per_cell_step': alg_bytes,
            'note': ('fused stencil + cell update; the cell update is FP64-'
                     'pipe- and issue-bound, so the HBM fraction is not '
                     'expected near 1 (DESIGN.md §4.1)'),
            # The binding ceiling: FP64-pipe instructions per cell-step of
            # this kernel (ncu, profiles/r01_opmix_c3.txt) against the FP64
            # FMA issue rate measured by mkb_measure_peaks on this pool
            # (profiles/r01_pipe_peaks.json).
            'fp64_pipe': fp64_pipe(n * n / world, kernel_ms),
        },
        'cpu_baseline': cpu,
    }
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
