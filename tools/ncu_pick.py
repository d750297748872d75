"""Prints the metrics that matter from an `ncu --page raw --csv` export: python tools/ncu_pick.py raw.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct', 'sm__warps_active.avg.pct', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'lts__t_sector_hit_rate.pct', 'launch__occupancy_limit', 'launch__grid_size', 'launch__registers_per_thread',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.max', 'smsp__average_warps_issue_stalled', 'launch__waves',
        'smsp__thread_inst_executed_per_inst_executed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__pipe_tensor', 'sass__inst_executed_local']
skip = ['pct_of_peak_sustained_elapsed', 'per_second', 'peak_sustained', 'not_issued']
for vals in rows[2:]:
    print('=====')
    for h, u, v in zip(hdr, units, vals):
        if any(k in h for k in want) and not any(k in h for k in skip):
            try:
                if 'stalled' in h and float(v) < 0.3:
                    continue
            except ValueError:
                pass
            print('%-90s %-12s %s' % (h, u, v))
