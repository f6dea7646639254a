"""Summarise an `ncu --page raw --csv` dump: key throughput counters and the top warp-stall reasons per kernel."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed.sum', 'lts__t_sectors_op_red.sum', 'lts__t_bytes.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_elapsed.max', 'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 'launch__registers_per_thread',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active']
for r in rows[2:]:
    print('----')
    for k in keys:
        for i, h in enumerate(hdr):
            if h == k:
                print(k, r[i], units[i])
    st = [(float(r[i].replace(',', '')), h) for i, h in enumerate(hdr)
          if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio') and r[i]]
    for v, h in sorted(st, reverse=True)[:10]:
        print('  %.2f %s' % (v, h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
