"""
Turns what scripts/collect_profiles.sh left in gpurun_out/ into the tracked
summaries under profiles/ (run here, no GPU needed; needs `ncu` to read reports).

    python scripts/summarize_profiles.py r01
"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys
from collections import Counter, OrderedDict

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(R, 'gpurun_out', os.environ.get('MKB_PROFILE_DIR', ''))
P = os.path.join(R, 'profiles')
tag = sys.argv[1] if len(sys.argv) > 1 else 'r01'

RAW = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size',
    'launch__registers_per_thread', 'launch__occupancy_limit_registers',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.sum', 'lts__t_sector_hit_rate.pct',
    'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum',
    'l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum',
    'gpc__cycles_elapsed.avg.per_second',
]


def ncu_csv(rep, page):
    out = subprocess.run(['ncu', '-i', rep, '--page', page, '--csv'],
                         capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def raw_metrics(rep):
    rows = ncu_csv(rep, 'raw')
    hdr, units, row = rows[0], rows[1], rows[2]
    out = OrderedDict()
    out['kernel'] = row[hdr.index('Kernel Name')]
    for m in RAW:
        if m in hdr:
            out[m] = (row[hdr.index(m)], units[hdr.index(m)])
    return out


def opmix(rep):
    rows = ncu_csv(rep, 'source')
    hdr = rows[1]
    i_src, i_ex, i_smp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
    stall_cols = [(i, h) for i, h in enumerate(hdr)
                  if h.startswith('stall_') and 'Not Issued' not in h]
    ops, smp, stalls = Counter(), Counter(), Counter()
    total, nwarps = 0, None
    for r in rows[2:]:
        if r and r[0] == 'Kernel Name':
            break
        if len(r) <= i_ex or not r[i_ex].isdigit():
            continue
        parts = r[i_src].split()
        if not parts:
            continue
        op = parts[1] if parts[0].startswith('@') else parts[0]
        op = op.rstrip(';')
        base = '.'.join(op.split('.')[:2]) if op.startswith(('MUFU', 'LDG', 'STG', 'LDL', 'STL', 'LDS', 'STS')) else op.split('.')[0]
        n = int(r[i_ex])
        if nwarps is None:
            nwarps = n
        ops[base] += n
        smp[base] += int(r[i_smp] or 0)
        total += n
        for i, h in stall_cols:
            stalls[h] += int(r[i] or 0)
    lines = ['warp-instructions executed: %d  (%.1f per warp = per cell-step)' % (total, total / nwarps)]
    ts = max(sum(smp.values()), 1)
    for op, n in ops.most_common(24):
        lines.append('  %-14s %8.1f /warp  %5.1f%% of instructions  %5.1f%% of samples'
                     % (op, n / nwarps, 100.0 * n / total, 100.0 * smp[op] / ts))
    lines.append('stall reasons (all samples):')
    tt = max(sum(stalls.values()), 1)
    for h, n in stalls.most_common(8):
        lines.append('  %-26s %5.1f%%' % (h, 100.0 * n / tt))
    return '\n'.join(lines), total / nwarps, ops, nwarps


def main():
    os.makedirs(P, exist_ok=True)
    md = ['# Profile summary %s' % tag, '',
          'Produced by `scripts/collect_profiles.sh` on a B200 (gpurun) and',
          '`scripts/summarize_profiles.py` here. Kernel times under ncu are cold-cache and',
          'serialised; bench numbers come from CUDA events in un-profiled runs.', '']
    for name in ('pipe_peaks.json', 'bench_n1.json', 'bench_ref.json', 'configs.txt'):
        src = os.path.join(G, name)
        if os.path.isfile(src):
            shutil.copy(src, os.path.join(P, '%s_%s' % (tag, name)))
    peaks = None
    pp = os.path.join(G, 'pipe_peaks.json')
    if os.path.isfile(pp):
        peaks = json.load(open(pp))
        md += ['## Measured pipe peaks (mkb_measure_peaks)', '',
               '| quantity | value |', '|---|---|']
        for k in ('fp64_fma_ginstr_s', 'fp32_fma_ginstr_s', 'mufu_ex2_gop_s', 'copy_gb_s', 'sm_clock_mhz', 'sm_count'):
            md.append('| %s | %.1f |' % (k, peaks[k]))
        md.append('')
    # launch list: share of the step kernel
    lp = os.path.join(G, 'launches.csv')
    if os.path.isfile(lp):
        rows = [r for r in csv.reader(open(lp)) if len(r) > 10 and r[0].isdigit()]
        tot = Counter()
        cnt = Counter()
        for r in rows:
            name = r[4].split('(')[0][:60]
            tot[name] += float(r[-1])
            cnt[name] += 1
        alln = sum(tot.values())
        with open(os.path.join(P, '%s_launches.csv' % tag), 'w') as f:
            f.write('kernel,launches,total_ns,share\n')
            for k, v in tot.most_common():
                f.write('"%s",%d,%.0f,%.4f\n' % (k, cnt[k], v, v / alln))
        md += ['## Launch list of `bench.py --steps 20 --warmup 3` under ncu (`%s_launches.csv`)' % tag, '',
               '| kernel | launches | total ms | share |', '|---|---|---|---|']
        for k, v in tot.most_common(6):
            md.append('| `%s` | %d | %.3f | %.1f%% |' % (k, cnt[k], v / 1e6, 100 * v / alln))
        md.append('')
    for w, title in (('c3', 'C3: decker-2009 fp64 RL, 2048x2048, conductance + scalar fields'),
                     ('lr91_fp32', 'LR1991 fp32 FE, 4096x4096 (C2/C4 kernel)'),
                     ('stencil32', 'stencil-only fp32, 8192x4096'),
                     ('stencil64', 'stencil-only fp64, 8192x4096')):
        rep = os.path.join(G, 'prof_%s.ncu-rep' % w)
        if not os.path.isfile(rep):
            continue
        m = raw_metrics(rep)
        text, per_warp, ops, nwarps = opmix(rep)
        with open(os.path.join(P, '%s_opmix_%s.txt' % (tag, w)), 'w') as f:
            f.write(title + '\n' + text + '\n')
        md += ['## ' + title, '', '`ncu --set full --clock-control none`, one launch of `mkb_cell_step`:', '',
               '| metric | value |', '|---|---|']
        for k, v in m.items():
            if k != 'kernel':
                md.append('| %s | %s %s |' % (k, v[0], v[1]))
        try:
            t = float(m['gpu__time_duration.sum'][0])
            unit = m['gpu__time_duration.sum'][1]
            t_ms = t if unit == 'ms' else (t / 1e3 if unit in ('us', 'usecond') else t / 1e6)
            rd = float(m['dram__bytes_read.sum'][0])
            wr = float(m['dram__bytes_write.sum'][0])
            scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}
            traffic = rd * scale[m['dram__bytes_read.sum'][1]] + wr * scale[m['dram__bytes_write.sum'][1]]
            md.append('| DRAM traffic per launch | %.3f GB |' % (traffic / 1e9))
            kf = os.path.join(G, 'key_%s.txt' % w)
            if w == 'c3' and os.path.isfile(kf):
                # what bench.py reads instead of literals (roofline.traffic, fp64_pipe)
                fp64 = sum(ops[k] for k in ('DFMA', 'DMUL', 'DADD', 'DSETP'))
                prof = {
                    'kernel_key': open(kf).read().strip(),
                    'workload': title,
                    'capture': 'ncu --set full --clock-control none, one launch (%s)' % tag,
                    'instr_per_cell_step': per_warp,
                    'fp64_instr_per_cell_step': fp64 / nwarps,
                    'dram_bytes_per_launch': traffic,
                    'duration_us_under_ncu': t_ms * 1e3,
                    'registers': int(float(m['launch__registers_per_thread'][0])),
                    'fp64_pipe_active_pct': float(m['sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'][0]),
                    'issue_active_pct': float(m['smsp__issue_active.avg.pct_of_peak_sustained_active'][0]),
                }
                with open(os.path.join(P, 'c3_kernel_profile.json'), 'w') as f:
                    json.dump(prof, f, indent=1)
            md.append('| DRAM GB/s under ncu | %.0f |' % (traffic / 1e9 / (t_ms * 1e-3)))
        except Exception:
            pass
        md += ['', 'Instruction mix (`%s_opmix_%s.txt`):' % (tag, w), '', '```', text, '```', '']
    with open(os.path.join(P, '%s_summary.md' % tag), 'w') as f:
        f.write('\n'.join(md) + '\n')
    print('\n'.join(md))


if __name__ == '__main__':
    main()
