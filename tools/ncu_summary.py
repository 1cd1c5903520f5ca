#!/usr/bin/env python3
"""Summarise an .ncu-rep (run where ncu is installed; no GPU needed):
  python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--sass] [--units-per-launch N]
Prints the raw-page metrics the roofline uses and, with --sass, per-SASS-region executed instructions and
stall samples of the first profiled kernel (needs -lineinfo + --import-source on)."""
import csv
import io
import re
import subprocess
import sys
from collections import Counter

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum',
        'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed']


def page(rep, name):
    out = subprocess.run(['ncu', '-i', rep, '--page', name, '--csv'], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    rows = page(rep, 'raw')
    hdr, units, data = rows[0], rows[1], rows[2:]
    for d in data:
        print('==', d[hdr.index('Kernel Name')])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f'  {w} = {d[i]} {units[i]}')
    if '--sass' not in sys.argv:
        return
    tiles = None
    if '--units-per-launch' in sys.argv:
        tiles = float(sys.argv[sys.argv.index('--units-per-launch') + 1]) / 32.0
    rows = page(rep, 'source')
    hdr = rows[1]
    isrc, iex, ith, isamp = (hdr.index(k) for k in ('Source', 'Instructions Executed', 'Avg. Threads Executed', '# Samples'))
    ins = []
    for r in rows[2:]:
        if r and r[0] == 'Kernel Name':
            break
        try:
            ins.append((r[isrc].strip(), int(r[iex]), float(r[ith]), int(r[isamp])))
        except (ValueError, IndexError):
            continue
    tot_ex = sum(i[1] for i in ins)
    tot_s = sum(i[3] for i in ins)
    scale = tiles or 1.0
    print(f'-- SASS regions of the first kernel: {len(ins)} instructions, {tot_ex} warp-instructions executed'
          + (f' = {tot_ex / tiles:.1f} per 32 units' if tiles else '') + f', {tot_s} stall samples')
    runs = []
    for k, (s, ex, th, sa) in enumerate(ins):
        op = s.split()[1] if s.startswith('@') else s.split()[0]
        if runs and runs[-1]['ex'] == ex and ex > 0:
            runs[-1]['n'] += 1
            runs[-1]['end'] = k
            runs[-1]['samp'] += sa
            runs[-1]['ops'].append(op)
            runs[-1]['hot'] = max(runs[-1]['hot'], (sa, s))
        else:
            runs.append({'ex': ex, 'n': 1, 'start': k, 'end': k, 'samp': sa, 'th': th, 'ops': [op], 'hot': (sa, s)})
    for r in runs:
        share = r['ex'] * r['n'] / max(tot_ex, 1)
        if share >= 0.004 or r['samp'] / max(tot_s, 1) >= 0.01:
            c = Counter(r['ops'])
            print(f"  [{r['start']:4d}-{r['end']:4d}] x{r['ex'] / scale:8.2f} n={r['n']:3d} inst%={100 * share:5.1f} thr={r['th']:4.1f} "
                  f"stall%={100 * r['samp'] / max(tot_s, 1):5.1f} {dict(c.most_common(5))} hot: {r['hot'][1][:60]} ({r['hot'][0]})")


if __name__ == '__main__':
    main()


def stall_totals(rep):
    """python -c 'import tools.ncu_summary as t; t.stall_totals(path)': warp-stall samples by reason, whole kernel."""
    rows = page(rep, 'source')
    hdr = rows[1]
    cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    tot = Counter()
    for r in rows[2:]:
        if r and r[0] == 'Kernel Name':
            break
        for i, h in cols:
            try:
                tot[h] += int(r[i])
            except (ValueError, IndexError):
                pass
    s = sum(tot.values())
    for h, v in tot.most_common(12):
        print(f'  {h:28s} {v:8d} {100 * v / max(s, 1):5.1f}%')
