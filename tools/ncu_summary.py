"""Prints the metrics that matter from an .ncu-rep (raw page) — run here, no GPU needed."""
import csv, subprocess, sys
rep = sys.argv[1]
pats = sys.argv[2:] or ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
    'launch__occupancy_limit', 'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_active',
    'smsp__issue_active.avg.pct', 'smsp__thread_inst_executed_per_inst_executed.ratio',
    'smsp__average_warps_issue_stalled', 'smsp__average_warp_latency_issue_stalled', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct', 'sm__cycles_elapsed.max',
    'smsp__cycles_active.avg', 'sm__pipe_tensor', 'lts__t_sector_hit_rate', 'gpc__cycles_elapsed.max',
    'smsp__warps_eligible.avg.per_cycle_active', 'smsp__pcsamp_warps_issue_stalled']
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print('== kernel', r[hdr.index('Kernel Name')][:60], 'grid', r[hdr.index('Grid Size')], 'block', r[hdr.index('Block Size')])
    for h, u, v in zip(hdr, units, r):
        name = h.split('.', 2)[-1] if h.count('.') >= 2 and h.split('.')[1] in ('TriageCompute',) else h
        if any(p in h for p in pats):
            print(f'  {h} [{u}] = {v}')
