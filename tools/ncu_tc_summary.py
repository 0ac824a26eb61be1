#!/usr/bin/env python3
"""Summary of an ncu --set full report for the TMA / tcgen05 kernels (tensor pipe, shared-memory operand fetch, L2, DRAM)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H, U = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_tensor.sum',
        'sm__inst_executed_pipe_uniform.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_bytes.sum.per_second', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes_read.sum.per_second', 'dram__bytes_write.sum.per_second',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'sm__cycles_elapsed.max', 'smsp__cycles_active.avg']
for V in rows[2:]:
    d = dict(zip(H, V))
    print('## kernel:', d.get('Kernel Name', '?')[:160])
    for k in want:
        for h in H:
            if h == k or h.endswith('.' + k):
                print('%-86s %-14s %s' % (k, U[H.index(h)], d[h]))
                break
    print()
