"""
Attributes the SASS of a generated kernel to the lines of its source (static
instruction counts, or executed counts / stall samples when an
`ncu --page source --print-source sass --csv` export of the same cubin is given).

    python scripts/attribute_sass.py kernel.cu kernel.cubin [ncu_source.csv]
"""
import csv
import re
import subprocess
import sys
from collections import Counter, defaultdict

FP64 = ('DFMA', 'DMUL', 'DADD', 'DSETP')


def sass_lines(cubin):
    out = subprocess.check_output(['nvdisasm', '-g', '-c', cubin]).decode()
    line = None
    func = None
    rows = []       # (func, address, opcode, source line)
    for text in out.splitlines():
        m = re.match(r'\s*\.text\.(\S+):', text)
        if m:
            func = m.group(1)
            continue
        m = re.match(r'\s*//## File "[^"]*", line (\d+)', text)
        if m:
            line = int(m.group(1))
            continue
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)', text)
        if m:
            rows.append((func, int(m.group(1), 16), m.group(3), line))
    return rows


def main():
    src = open(sys.argv[1]).read().splitlines()
    rows = sass_lines(sys.argv[2])
    executed = None
    if len(sys.argv) > 3:
        data = list(csv.reader(open(sys.argv[3])))[2:]
        executed = [int(r[5]) for r in data]
        samples = [int(r[4]) for r in data]
    main_rows = [r for r in rows if r[0] == rows[0][0]]
    if executed is not None and len(executed) != len(rows):
        print('warning: %d SASS rows, %d csv rows' % (len(rows), len(executed)))
    by_line = defaultdict(Counter)
    for i, (func, addr, op, line) in enumerate(rows):
        w = 1 if executed is None else executed[i] / 131072.0
        base = op.split('.')[0]
        key = line if func == rows[0][0] else func
        by_line[key]['all'] += w
        if base in FP64:
            by_line[key]['fp64'] += w
        if executed is not None:
            by_line[key]['samples'] += samples[i]
    tot = sum(c['all'] for c in by_line.values())
    tot64 = sum(c['fp64'] for c in by_line.values())
    print('total %.0f instructions, %.0f FP64' % (tot, tot64))
    for key, c in sorted(by_line.items(), key=lambda kv: -kv[1]['all'])[:int(sys.argv[4]) if len(sys.argv) > 4 else 60]:
        text = src[key - 1].strip()[:110] if isinstance(key, int) and key <= len(src) else str(key)
        print('%7.1f all %7.1f fp64 %s | %s' % (c['all'], c['fp64'],
              ('%5d smp' % c['samples']) if executed is not None else '', text))


if __name__ == '__main__':
    main()
