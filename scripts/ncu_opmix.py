"""Aggregates an `ncu --page source --csv` dump by SASS opcode and stall reason."""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
i_src = hdr.index('Source')
i_ex = hdr.index('Instructions Executed')
i_smp = hdr.index('# Samples')
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
ops, smp = Counter(), Counter()
stalls = Counter()
total = 0
nwarps = None
for r in rows[2:]:
    if r and r[0] == 'Kernel Name':
        break               # only the first captured launch
    if len(r) <= i_ex or not r[i_ex].isdigit():
        continue
    src = r[i_src].strip()
    parts = src.split()
    if not parts:
        continue
    op = parts[0]
    if op.startswith('@'):
        op = parts[1]
    op = op.rstrip(';')
    base = '.'.join(op.split('.')[:2]) if op.startswith(('MUFU', 'F2F', 'I2F', 'F2I', 'DSETP', 'LDG', 'STG', 'LDL', 'STL')) else op.split('.')[0]
    n = int(r[i_ex] or 0)
    if nwarps is None:
        nwarps = n
    ops[base] += n
    smp[base] += int(r[i_smp] or 0)
    total += n
    for i, h in stall_cols:
        stalls[h] += int(r[i] or 0)
print('total warp-instructions', total, 'per warp', total / nwarps)
tot_s = sum(smp.values())
for op, n in ops.most_common(40):
    print('%-14s %12d  %7.1f/warp  %5.1f%% inst  %5.1f%% samples' % (op, n, n / nwarps, 100.0 * n / total, 100.0 * smp[op] / max(tot_s, 1)))
print('stalls:')
ts = sum(stalls.values())
for h, n in stalls.most_common(12):
    print('  %-28s %5.1f%%' % (h, 100.0 * n / max(ts, 1)))
