#!/usr/bin/env python
"""Print the headline numbers of an `ncu --page raw --csv` export."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
d = dict(zip(rows[0], rows[2]))
keys = ['Kernel Name', 'gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.per_cycle_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'smsp__warps_active.avg.per_cycle_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps']
for k in keys:
    print('%-60s %s' % (k, d.get(k)))
st = []
for k, v in d.items():
    if k.startswith('smsp__pcsamp_warps_issue_stalled_') and not k.endswith('_not_issued'):
        try: st.append((float(v), k.replace('smsp__pcsamp_warps_issue_stalled_', '')))
        except ValueError: pass
tot = sum(v for v, _ in st) or 1
print('stall samples:', ', '.join('%s %.0f%%' % (k, 100 * v / tot) for v, k in sorted(st, reverse=True)[:9]))
