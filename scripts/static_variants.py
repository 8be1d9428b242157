"""Static evidence (ptxas -v, SASS opcode counts) for the kernel variants; no GPU needed.

    python scripts/static_variants.py > profiles/r01_static_variants.md
"""
import os, re, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
sys.path.insert(0, os.path.join(R, 'scripts'))
import myokit_b200, myokit
from myokit_b200 import workloads, multigpu
from sass_stats import sass_counts


def row(label, src, func='mkb_cell_step'):
    pf, log = sass_counts(src)
    c = pf[func]
    tot = sum(c.values())
    fp64 = c['DFMA'] + c['DADD'] + c['DMUL'] + c['DSETP']
    fp32 = c['FFMA'] + c['FADD'] + c['FMUL']
    mufu = sum(v for k, v in c.items() if k.startswith('MUFU'))
    m = re.search(r"Compiling entry function '%s'.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers" % func, log, re.S)
    st, ss, sl, regs = m.groups() if m else ('?',) * 4
    print('| %s | %d | %d | %d | %d | %s | %s / %s |' % (label, tot, fp64, fp32, mufu, regs, ss, sl))


print('# Static comparison of kernel variants (round 1, no GPU: `scripts/static_variants.py`)')
print()
print('SASS instruction counts of `mkb_cell_step` as compiled by NVRTC for sm_100a (static: both sides of branches and')
print('library slow paths are included), registers and spill bytes per thread from `ptxas -v`. Timings belong to round 2.')
print()
print('## C3 kernel: decker-2009 fp64 Rush-Larsen, conductance fields + one scalar field')
print()
print('| variant | instructions | FP64 pipe | FP32 | MUFU | registers | spill st / ld (B) |')
print('|---|---|---|---|---|---|---|')
for label, opts in (('default', {}), ('`const_div=False` (before this round\'s rewrite)', dict(const_div=False)),
                    ('`div_parallel=True`', dict(div_parallel=True)), ('`fast_exp=\'estrin\'`', dict(fast_exp='estrin')),
                    ('estrin + div_parallel', dict(fast_exp='estrin', div_parallel=True)),
                    ('`fast_exp=\'table\'`', dict(fast_exp='table')),
                    ('libdevice arithmetic (`fast_div/fast_exp/pow_multiply/const_div` off)',
                     dict(fast_div=False, fast_exp=False, pow_multiply=False, const_div=False))):
    s = workloads.c3_hetero(myokit_b200.SimulationCUDA, nx=64)
    s.set_kernel_options(**opts)
    row(label, s.kernel_source())
s = workloads.c3_hetero(myokit_b200.SimulationCUDA, nx=64)
s.set_kernel_options(split_gates=True)
row('`split_gates=True`: `mkb_cell_step`', s.kernel_source())
row('`split_gates=True`: `mkb_gate_step` (10 gating variables)', s.kernel_source(), 'mkb_gate_step')
box = {}


def work(comm):
    for lean in (False, True):
        s = workloads.c3_hetero(myokit_b200.SimulationCUDA, nx=64, comm=comm)
        s.set_kernel_options(slab_lean=lean)
        if comm.rank == 0:
            box[lean] = s.kernel_source()


multigpu.run_threads(2, work)
row('row-slab kernel (multi-GPU), default form', box[False])
row('row-slab kernel, `slab_lean=True`', box[True])
print()
print('## LR1991 fp32 forward Euler (C2 / C4 kernel), homogeneous grid')
print()
print('| variant | instructions | FP64 pipe | FP32 | MUFU | registers | spill st / ld (B) |')
print('|---|---|---|---|---|---|---|')
for label, opts in (('default (libdevice `expf`)', {}), ('`fast_exp=\'ex2\'`', dict(fast_exp='ex2'))):
    s = workloads.c2_planar(myokit_b200.SimulationCUDA, n=64)
    s.set_kernel_options(**opts)
    row(label, s.kernel_source())
s = workloads.c2_planar(myokit_b200.SimulationCUDA, n=64)
s = myokit_b200.SimulationCUDA(s._model, s._protocol, ncells=(64, 64), precision=myokit.SINGLE_PRECISION, native_maths=True)
row('`native_maths=True` (`__expf`, as the reference\'s `native_exp`)', s.kernel_source())
