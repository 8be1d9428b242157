"""Host-side overhead of successive run_fields() calls on a resident session (diagnostic)."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import myokit_b200
from myokit_b200 import workloads

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
s = workloads.c3_hetero(myokit_b200.SimulationCUDA, nx=n)
t0 = time.perf_counter()
s.run_fields(200 * 0.005, ['membrane.V'], log_interval=1.0)
print('cold call %.3f s' % (time.perf_counter() - t0), s.last_run_info()['host_seconds'])
for k in (20, 20, 20, 200, 200, 1000, 1000, 20):
    t0 = time.perf_counter()
    tt, f = s.run_fields(k * 0.005, ['membrane.V'], log_interval=1.0)
    dt = time.perf_counter() - t0
    i = s.last_run_info()
    hs = i['host_seconds']
    print('K=%4d  total %.4f s  device %.4f s  arm %.4f steps %.4f collect %.4f  rows %d  -> %.3e cell-steps/s'
          % (k, dt, i['device_ms'] * 1e-3, hs['arm'], hs['steps'], hs['collect'], len(tt), n * n * i['steps'] / dt))
