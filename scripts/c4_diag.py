import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import numpy as np
import myokit_b200, myokit
from oracle.oracle import OracleSimulation
SP = myokit.SINGLE_PRECISION
n = 512
m, _, _ = myokit.load('example')
ns = m.count_states()
init = np.array(m.initial_values(True))
state = np.tile(init, n * n).reshape(n, n, ns)
iv = m.get('membrane.V').index()
ih, ij = m.get('ina.h').index(), m.get('ina.j').index()
state[:n // 2, 40:60, iv] = 10.0
state[:n // 2, 0:40, ih] = 0.0
state[:n // 2, 0:40, ij] = 0.0
state[:n // 2, 0:40, iv] = -55.0
def run(cls, kw, run_kw, opts=None, dur=20):
    s = cls(m, None, ncells=(n, n), precision=SP, **kw)
    s.set_conductance(1, 1); s.set_paced_cells(0, 0, 0, 0); s.set_step_size(0.005)
    s.set_state(state.ravel())
    if opts: s.set_kernel_options(**opts)
    r = s.run(dur, log=['engine.time'], log_interval=1, **run_kw)
    return np.asarray(r[1]) if isinstance(r, tuple) else s.state_array()
for dur in (5, 10, 20):
    sb = run(OracleSimulation, dict(openmp=True), dict(nthreads=os.cpu_count()), dur=dur)
    scale = np.abs(sb).reshape(-1, ns).max(axis=0)
    for name, opts in (('default', None), ('ieee div', dict(fast_div=False)), ('ieee div nofmad', dict(fast_div=False, fmad=False))):
        sa = run(myokit_b200.SimulationCUDA, {}, {}, opts, dur=dur)
        print('   nan: cuda', int(np.isnan(sa).sum()), 'oracle', int(np.isnan(sb).sum()), 'scale', scale)
        rel = np.abs(sa - sb).reshape(-1, ns) / scale
        k = np.unravel_index(np.argmax(rel), rel.shape)
        print(dur, name, 'rel max %.3g' % rel.max(), 'state', k[1], 'cell', divmod(int(k[0]), n), 'p99.99 %.3g' % np.quantile(rel, 0.9999), flush=True)
