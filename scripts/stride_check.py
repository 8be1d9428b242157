import sys, os
sys.path.insert(0, os.getcwd())
import myokit_b200
from myokit_b200 import workloads
S = myokit_b200.SimulationCUDA
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
for steps in (30, 300):
    s = workloads.c3_hetero(S, nx=n)
    i = s.benchmark_steps(steps if n <= 4096 else max(steps // 10, 10), warmup=5)
    print('MKB_STRIDE_PAD=%s %d^2 %d steps: %.4f ms/step' % (os.environ.get('MKB_STRIDE_PAD', '0'), n, i['steps'], i['device_ms'] / i['steps']), flush=True)
