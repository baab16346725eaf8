"""Key metrics per kernel from `ncu -i X.ncu-rep --page raw --csv` on stdin."""
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed.avg.per_cycle_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.avg',
        'local_load', 'smsp__inst_executed_op_local']
for r in rows[2:]:
    print('----', r[hdr.index('Kernel Name')])
    for k in keys:
        for i, h in enumerate(hdr):
            if h == k or (k in ('local_load',) and k in h):
                print('   %-70s %s %s' % (h, r[i], units[i]))
    items = []
    for i, h in enumerate(hdr):
        if h.startswith('smsp__pcsamp_warps_issue_stalled') and not h.endswith('not_issued'):
            try:
                items.append((float(r[i]), h.replace('smsp__pcsamp_warps_issue_stalled_', '')))
            except ValueError:
                pass
    tot = sum(v for v, _ in items) or 1
    print('   stall samples:', ', '.join('%s %.0f%%' % (h, 100 * v / tot) for v, h in sorted(items, reverse=True)[:9]))
