"""Print the handful of metrics the profile summaries quote from an `ncu -i X --page raw --csv` dump (stdin or file)."""
import csv, sys
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'launch__registers_per_thread',
        'launch__grid_size', 'sm__cycles_active.avg', 'launch__shared_mem_per_block_dynamic',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed.avg.per_cycle_active']
rows = list(csv.reader(open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print('---', r[idx['Kernel Name']][:60])
    for w in WANT:
        if w in idx:
            print('  %-84s %14s %s' % (w, r[idx[w]], rows[1][idx[w]]))
