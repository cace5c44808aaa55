"""Print the key metrics of an `ncu --page raw --csv` dump, one column per profiled launch."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h, units, data = rows[0], rows[1], rows[2:]
want = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__warps_eligible.avg.per_cycle_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum',
        'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_fmaheavy.sum', 'sm__inst_executed_pipe_fp64.sum']
idx = {n: i for i, n in enumerate(h)}
for w in want:
    if w in idx:
        print(f"{w:62s} {units[idx[w]]:10s} " + " | ".join(f"{r[idx[w]]:>14s}" for r in data))
print("-- stall reasons (warp cycles per issued instruction), > 0.3 --")
for n, i in sorted(idx.items()):
    if n.startswith('smsp__average_warps_issue_stalled') and n.endswith('per_issue_active.ratio'):
        vals = [float(r[i].replace(',', '')) for r in data]
        if max(vals) > 0.3:
            print(f"{n[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:40s} " + " | ".join(f"{v:8.2f}" for v in vals))
