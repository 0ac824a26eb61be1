#!/usr/bin/env python3
"""Summarise ncu outputs brought back in gpurun_out/ into small text files for profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv            > profiles/<tag>_launches.txt
    python tools/ncu_summary.py full gpurun_out/prof.ncu-rep [n_units]      > profiles/<tag>_full.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed',
        'smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed',
        'smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed',
        'sm__sass_thread_inst_executed_op_ffma_pred_on.sum.peak_sustained',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes_read.sum.per_second',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.max',
        'smsp__sass_average_branch_targets_threads_uniform.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    H, data = rows[hi], rows[hi + 1:]
    ki, vi = H.index('Kernel Name'), H.index('Metric Value')
    agg = collections.OrderedDict()
    for r in data:
        if len(r) > vi:
            agg.setdefault(r[ki], []).append(float(r[vi].replace(',', '')))
    tot = sum(sum(v) for v in agg.values())
    print('# ncu --metrics gpu__time_duration.sum --clock-control none  (cold-cache, serialised: compare SHARES)')
    print('%-100s %5s %12s %7s' % ('kernel', 'n', 'total_ms', 'share'))
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print('%-100s %5d %12.3f %7.3f' % (k[:100], len(v), sum(v) / 1e6, sum(v) / tot))


def full(path, n_units=None):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    H, U = rows[0], rows[1]
    for V in rows[2:]:
        d = dict(zip(H, V))
        print('## kernel:', d.get('Kernel Name', '?'))
        for k in KEYS:
            if k in d:
                print('%-80s %-16s %s' % (k, U[H.index(k)], d[k]))
        if n_units and 'smsp__inst_executed.sum' in d:
            print('warp instructions per unit (voxel): %.1f' % (float(d['smsp__inst_executed.sum'].replace(',', '')) / n_units))
    src = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    his = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
    if his:
        H, D = rows[his[0]], rows[his[0] + 1:]
        ie, isrc = H.index('Instructions Executed'), H.index('Source')
        stall_cols = [i for i, h in enumerate(H) if h.startswith('stall_') and 'Not Issued' not in h]
        mix, stalls = collections.Counter(), collections.Counter()
        for r in D:
            if len(r) <= ie or not r[ie].isdigit():
                continue
            op = r[isrc].split()
            o = op[1] if op[0].startswith('@') else op[0]
            mix[o.split('.')[0]] += int(r[ie])
            for i in stall_cols:
                stalls[H[i]] += int(r[i] or 0)
        tot = sum(mix.values())
        print('## SASS opcode mix (executed warp instructions)')
        for k, v in mix.most_common(16):
            print('  %-10s %6.2f%%' % (k, 100.0 * v / tot))
        ts = sum(stalls.values())
        print('## warp stall samples')
        for k, v in stalls.most_common(8):
            print('  %-28s %6.2f%%' % (k, 100.0 * v / max(ts, 1)))


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2])
    else:
        full(sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else None)
