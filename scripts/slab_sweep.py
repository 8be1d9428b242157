"""Kernel options of the row-slab form, timed on 2 slabs (two threads, two devices, direct peer access)."""
import sys, os
sys.path.insert(0, os.getcwd())
import myokit_b200
from myokit_b200 import workloads, multigpu, capi
ny = int(sys.argv[1]) if len(sys.argv) > 1 else 512          # rows of the whole grid: two slabs of ny / 2
variants = [('default', {}), ('no overlap', dict(overlap=False)), ('overlap', dict(overlap=True)),
            ('no stage', dict(stage=False)), ('no stage, no overlap', dict(stage=False, overlap=False)),
            ('no stage, overlap', dict(stage=False, overlap=True)),
            ('direct stores', dict(stage_store=False)), ('direct stores, no overlap', dict(stage_store=False, overlap=False)),
            ('one group la8', dict(stage_group=64, load_ahead=8)),
            ('not lean', dict(slab_lean=False)), ('256x1', dict(block=(256, 1)))]
ndev = capi.device_count()
def work(comm):
    out = []
    for name, o in variants:
        s = workloads.c3_hetero(myokit_b200.SimulationCUDA, nx=2048, ny=ny, device=comm.rank % ndev, comm=comm)
        s.set_kernel_options(**o)
        i = s.benchmark_steps(300, warmup=70)
        ms = max(comm.allgather(i['device_ms'] / i['steps']))
        out.append((name, ms))
    return out
res = multigpu.run_threads(2, work)[0]
s = workloads.c3_hetero(myokit_b200.SimulationCUDA, nx=2048, ny=ny // 2)
i = s.benchmark_steps(300, warmup=70)
print('2048 x %d unsharded on one GPU: %.4f ms/step' % (ny // 2, i['device_ms'] / i['steps']))
for name, ms in res:
    print('2 slabs of 2048 x %d, %-12s %.4f ms/step' % (ny // 2, name, ms), flush=True)
