#!/usr/bin/env python
"""One-screen summary of an ncu report (first kernel matching a substring): time, DRAM bytes, instruction count, issue /
occupancy / pipe utilisation, shared-memory wavefronts and the main stall reasons.  usage: ncu_summary.py <rep> [kernel-substr]"""
import csv, subprocess, sys
rep = sys.argv[1]
sub = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'lts__t_sectors_srcunit_tex_op_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active']
for r in rows[2:]:
    d = dict(zip(h, r))
    if sub and sub not in d.get('Kernel Name', ''):
        continue
    print("kernel:", d.get('Kernel Name'))
    for k in keys:
        if k in d:
            print("  %-78s %s %s" % (k, d[k], units[h.index(k)]))
    for i, k in enumerate(h):
        if 'issue_stalled' in k and k.endswith('.ratio') and 'not_issued' not in k:
            try:
                v = float(r[i])
            except ValueError:
                continue
            if v >= 0.2:
                print("  stall %-72s %.2f" % (k.split('issue_stalled_')[1].replace('.ratio', ''), v))
    break
