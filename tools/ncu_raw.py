import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[0]; units=rows[1]; vals=rows[2]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','dram__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed','sm__cycles_active.avg','smsp__cycles_active.avg','sm__cycles_elapsed.max','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__grid_size','lts__t_sectors_srcunit_tex_op_read.sum','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_active']
for i,h in enumerate(hdr):
    if h in want: print(h, units[i], vals[i])
    if 'smsp__average_warp' in h and 'issue_stalled' in h and 'ratio' in h and float(vals[i] or 0)>0.15: print(h.replace('smsp__average_warps_issue_stalled_','stall_'), vals[i])
