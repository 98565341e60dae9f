#!/usr/bin/env python
"""Key metrics of every launch of an ncu report, one column per launch (a superset of tools/ncu_summary.py's table: adds the
stall reasons per issue, shared-memory wavefronts / bank conflicts and the executed warp instructions).  tools/ncu_keys.py rep.ncu-rep"""
import csv, subprocess, sys
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
h, u, data = rows[0], rows[1], rows[2:]
want = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active']
ki = h.index("Kernel Name")
print("kernel", " | ".join(r[ki].split("(")[0][-40:] for r in data))
for k in h:
    if k in want or ('issue_stalled' in k and k.endswith('per_issue_active.ratio')):
        i = h.index(k)
        vals = [r[i] for r in data]
        if 'issue_stalled' in k and all(float(v or 0) < 0.2 for v in vals):
            continue
        print(k.replace('smsp__average_warps_issue_stalled_', 'stall_').replace('_per_issue_active.ratio', ''), '[%s]' % u[i], " | ".join(vals))
