import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import numpy as np
import myokit_b200, myokit
from myokit_b200 import workloads
from oracle.oracle import OracleSimulation
SP, DP = myokit.SINGLE_PRECISION, myokit.DOUBLE_PRECISION
nx, ny = 264, 77
for prec, hetero in ((SP, True), (SP, False)):
    def make(cls, **opts):
        s = workloads.stencil_only(cls, nx, ny, precision=prec, hetero=hetero)
        if opts:
            s.set_kernel_options(**opts)
        return s
    res = {}
    for name, opts in (('stream', dict(stream=True, fmad=False)), ('vector', dict(stream=False, fmad=False)),
                       ('scalar', dict(stream=False, fmad=False, cells_per_thread=1, rows_per_thread=1)),
                       ('stream2', dict(stream=True, fmad=False, use_graphs=False))):
        s = make(myokit_b200.SimulationCUDA, **opts)
        t, f = s.run_fields(6, ['membrane.V'], log_interval=0.5)
        res[name] = f['membrane.V']
    o = make(OracleSimulation)
    keys = ['%d.%d.membrane.V' % (x, y) for y in range(ny) for x in range(nx)]
    ol, ostate = o.run(6, log=['engine.time'] + keys, log_interval=0.5)
    want = np.array([[ol[k][i] for k in keys] for i in range(len(ol['engine.time']))]).reshape(-1, ny, nx)
    for name, got in res.items():
        bad = np.argwhere(got != want.astype(got.dtype))
        print(prec, hetero, name, 'mismatches', len(bad), 'first', bad[:5].tolist(), 'max abs', np.abs(got - want).max() if len(bad) else 0, flush=True)
