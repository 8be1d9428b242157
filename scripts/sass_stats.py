"""Static SASS opcode counts of the generated kernel for a set of options (no GPU)."""
import os, sys, re, subprocess, tempfile
from collections import Counter
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import myokit_b200, myokit
from myokit_b200 import workloads, capi


def sass_counts(src):
    cubin, log = capi.jit_compile(src.code, src.options + ('--ptxas-options=-v',))
    with tempfile.NamedTemporaryFile(suffix='.cubin') as f:
        f.write(cubin); f.flush()
        out = subprocess.check_output(['cuobjdump', '-sass', f.name]).decode()
    ops = Counter()
    func = None
    per_func = {}
    for line in out.splitlines():
        m = re.match(r'\s+Function : (\S+)', line)
        if m:
            func = m.group(1); per_func[func] = Counter(); continue
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if m and func:
            op = m.group(2)
            base = '.'.join(op.split('.')[:2]) if op.startswith(('MUFU',)) else op.split('.')[0]
            per_func[func][base] += 1
    regs = re.search(r"Used (\d+) registers", log)
    return per_func, log


if __name__ == '__main__':
    import ast
    opts = eval(sys.argv[1]) if len(sys.argv) > 1 else {}
    s = workloads.c3_hetero(myokit_b200.SimulationCUDA, nx=16)
    s.set_kernel_options(**opts)
    per_func, log = sass_counts(s.kernel_source())
    for func, c in per_func.items():
        tot = sum(c.values())
        fp64 = c['DFMA'] + c['DADD'] + c['DMUL'] + c['DSETP']
        print(func, 'total', tot, 'fp64', fp64, 'code KB', tot * 16 // 1024)
        print('   ', ', '.join('%s %d' % kv for kv in c.most_common(16)))
    for l in log.splitlines():
        if 'registers' in l or 'spill' in l:
            print(l.strip())
