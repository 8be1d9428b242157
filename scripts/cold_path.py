"""Where the first run of a new simulation spends its time (MKB_DEBUG_TIMING phases)."""
import os, sys, time
os.environ['MKB_DEBUG_TIMING'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import myokit_b200
from myokit_b200 import workloads
import torch
torch.zeros(1, device='cuda')
for rep in range(2):
    s = workloads.c3_hetero(myokit_b200.SimulationCUDA, nx=2048)
    t0 = time.perf_counter()
    s.run_fields(200 * 0.005, ['membrane.V'], log_interval=1.0)
    print('rep %d: cold call %.3f s' % (rep, time.perf_counter() - t0), s.last_run_info()['host_seconds'], flush=True)
    s.close()
