"""
How tightly ptxas packs dependent FP64 instructions in a generated kernel (no GPU):
for every DFMA / DMUL / DADD the distance, in instructions, to the producer of its
nearest source operand, and how many DFMAs take a constant from a uniform register.
A kernel whose FP64 instructions mostly follow their producers within two
instructions has few independent chains in flight (`stall_wait` in ncu).

    python scripts/sass_chains.py "dict()" "dict(tile_loop=True)" "dict(stage=False)"
"""
import os, re, statistics, subprocess, sys, tempfile
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import myokit_b200
from myokit_b200 import workloads, capi


def sass_of(opts):
    s = workloads.c3_hetero(myokit_b200.SimulationCUDA, nx=2048)
    if opts:
        s.set_kernel_options(**opts)
    src = s.kernel_source()
    cubin, log = capi.jit_compile(src.code, src.options)
    with tempfile.NamedTemporaryFile(suffix='.cubin') as f:
        f.write(cubin)
        f.flush()
        return subprocess.check_output(['cuobjdump', '-sass', f.name]).decode()


def analyse(text):
    last, gaps, k, ur, dfma, total = {}, [], 0, 0, 0, 0
    for line in text.splitlines():
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)\s+(.*?);', line)
        if not m:
            continue
        k += 1
        op, ops = m.group(2), m.group(3)
        if op != 'NOP':
            total += 1
        regs = re.findall(r'\bR(\d+)\b', ops)
        if op.split('.')[0] in ('DFMA', 'DMUL', 'DADD'):
            if op == 'DFMA':
                dfma += 1
                ur += 'UR' in ops
            srcs = [int(x) for x in regs[1:]]
            d = [k - last[x] for x in srcs + [x + 1 for x in srcs] if x in last]
            if d:
                gaps.append(min(d))
            if regs:
                last[int(regs[0])] = last[int(regs[0]) + 1] = k
        elif regs and not op.startswith(('ST', 'BRA', 'EXIT', 'BAR')):
            last[int(regs[0])] = k
    n = len(gaps)
    return dict(instructions=total, fp64=n, median=statistics.median(gaps),
                within2=100.0 * sum(1 for x in gaps if x <= 2) / n,
                within4=100.0 * sum(1 for x in gaps if x <= 4) / n,
                dfma_ur=100.0 * ur / max(dfma, 1))


if __name__ == '__main__':
    for arg in sys.argv[1:] or ['dict()']:
        r = analyse(sass_of(eval(arg)))
        print('%-58s %5d instr, %4d FP64: producer <= 2 back %4.1f %%, <= 4 back %4.1f %% (median %g); '
              'DFMA with a uniform-register operand %4.1f %%' % (
                  arg, r['instructions'], r['fp64'], r['within2'], r['within4'], r['median'], r['dfma_ur']))
