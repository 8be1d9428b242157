"""Kernel-option sweep on the C3 workload: compile stats here, timings on a GPU."""
import os, sys, time, re
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import myokit_b200, myokit
from myokit_b200 import workloads, capi

grid = int(os.environ.get('SWEEP_GRID', '2048'))
steps = int(os.environ.get('SWEEP_STEPS', '20'))
variants = [
    ('default', dict()),
    # prepared in round 1 without GPU time left to measure them:
    ('div_parallel', dict(div_parallel=True)),
    ('exp estrin', dict(fast_exp='estrin')),
    ('estrin + div_parallel', dict(fast_exp='estrin', div_parallel=True)),
    ('exp table (fewest FP64 instructions)', dict(fast_exp='table')),
    ('table + div_parallel', dict(fast_exp='table', div_parallel=True)),
    ('split gates (two kernels per step)', dict(split_gates=True)),
    ('split gates + div_parallel', dict(split_gates=True, div_parallel=True)),
    ('const_div off', dict(const_div=False)),
]
only = os.environ.get('SWEEP_ONLY')
gpu = capi.device_count() > 0
for name, opts in variants:
    if only and only not in name:
        continue
    s = workloads.c3_hetero(myokit_b200.SimulationCUDA, nx=grid if gpu else 16)
    s.set_kernel_options(**opts)
    src = s.kernel_source()
    t0 = time.time()
    cubin, log = capi.jit_compile(src.code, src.options + ('--ptxas-options=-v',))
    m = re.search(r"Compiling entry function 'mkb_cell_step'.*?Used (\d+) registers", log, re.S)
    sp = re.search(r"mkb_cell_step\n.*?(\d+) bytes stack frame, (\d+) bytes spill stores", log, re.S)
    line = '%-22s regs %s stack/spill %s cubin %d KB compile %.1fs' % (
        name, m.group(1) if m else '?', sp.groups() if sp else '?', len(cubin) // 1024, time.time() - t0)
    if gpu:
        info = s.benchmark_steps(steps, warmup=3)
        ms = info['device_ms'] / info['steps']
        line += '  %.3f ms/step  %.3e cell-steps/s' % (ms, grid * grid / ms * 1e3)
    print(line, flush=True)
