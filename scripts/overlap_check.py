"""Overlapping steps (kernelgen overlap=True, PDL + per-tile step counters): same bits, and what it buys by grid size."""
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
import myokit_b200, myokit
from myokit_b200 import workloads
S = myokit_b200.SimulationCUDA

def fields(nx, ny, dur, **opts):
    s = workloads.c3_hetero(S, nx=nx, ny=ny)
    s.set_kernel_options(**opts)
    t, f = s.run_fields(dur, ['membrane.V'], log_interval=dur / 4)
    return f['membrane.V'], s.state_array(), s.last_run_info()

for nx, ny in ((300, 70), (512, 512)):
    a = fields(nx, ny, 3.0, overlap=True)
    b = fields(nx, ny, 3.0, overlap=False)
    print('%d x %d: overlap == plain: V %s state %s, V range %.1f..%.1f, launches %d / %d' % (
        nx, ny, np.array_equal(a[0], b[0]), np.array_equal(a[1], b[1]), b[0].min(), b[0].max(),
        a[2]['kernel_launches'], b[2]['kernel_launches']), flush=True)
    # many short runs back to back (re-arm: counters restart)
s = workloads.c3_hetero(S, nx=256, ny=64); s.set_kernel_options(overlap=True)
r = workloads.c3_hetero(S, nx=256, ny=64); r.set_kernel_options(overlap=False)
for k in range(3):
    ta, fa = s.run_fields(1.2, ['membrane.V'], log_interval=0.4)
    tb, fb = r.run_fields(1.2, ['membrane.V'], log_interval=0.4)
    print('run %d equal %s' % (k, np.array_equal(fa['membrane.V'], fb['membrane.V'])), flush=True)
for nx, ny in ((2048, 2048), (2048, 1024), (2048, 512), (2048, 256), (1024, 1024), (512, 512), (256, 256)):
    out = []
    for ov in (False, True):
        s = workloads.c3_hetero(S, nx=nx, ny=ny)
        s.set_kernel_options(overlap=ov)
        i = s.benchmark_steps(200 if nx * ny > 1e6 else 1000, warmup=70)
        out.append(i['device_ms'] / i['steps'])
    print('%4d x %4d: plain %.4f ms/step, overlap %.4f ms/step (%.1f %%)' % (nx, ny, out[0], out[1], 100 * (out[1] / out[0] - 1)), flush=True)
# fp32 LR1991 (C2 / C4 kernel)
for n in (512, 2048):
    out = []
    for ov in (False, True):
        s = workloads.c2_planar(S, n)
        s.set_kernel_options(overlap=ov)
        i = s.benchmark_steps(2000 if n <= 512 else 300, warmup=70)
        out.append(i['device_ms'] / i['steps'])
    print('LR91 fp32 %d^2: plain %.4f ms/step, overlap %.4f ms/step (%.1f %%)' % (n, out[0], out[1], 100 * (out[1] / out[0] - 1)), flush=True)
