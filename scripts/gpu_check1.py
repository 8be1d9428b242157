import sys, time
import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R,'tests'))
import numpy as np
import util, myokit
from util import run_pair, max_abs_diff
m, p = util.example()
for fmad in (True, False):
    t=time.time()
    cl, cs, ol, os_ = run_pair(m, p, 128, 100, ['engine.time','membrane.V','engine.pace'], 1.0, cfg=dict(conductance=(10,), paced_cells=(5,)), fmad=fmad)
    print('1d fp64 fmad', fmad, 'rows', len(cl['engine.time']), 'dV', max_abs_diff(cl, ol, suffix='membrane.V'), 'dt', max_abs_diff(cl, ol, ['engine.time','engine.pace']), 'dstate', np.max(np.abs(cs-os_)), time.time()-t)
t=time.time()
cl, cs, ol, os_ = run_pair(m, p, (32,16), 20, ['engine.time','membrane.V','membrane.i_diff','ica.ICa'], 0.5, cfg=dict(conductance=(10,5), paced_cells=(3,16,0,0)))
print('2d fp64', 'rows', len(cl['engine.time']), 'dV', max_abs_diff(cl, ol, suffix='membrane.V'), 'didiff', max_abs_diff(cl, ol, suffix='i_diff'), 'dICa', max_abs_diff(cl, ol, suffix='ICa'), 'dstate', np.max(np.abs(cs-os_)), time.time()-t)
cl, cs, ol, os_ = run_pair(m, p, (32,16), 20, ['engine.time','membrane.V'], 0.5, precision=myokit.SINGLE_PRECISION, cfg=dict(conductance=(10,5), paced_cells=(3,16,0,0)))
print('2d fp32', 'rows', len(cl['engine.time']), 'dV', max_abs_diff(cl, ol, suffix='membrane.V'), 'dstate', np.max(np.abs(cs-os_)))
