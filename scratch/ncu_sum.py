import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
h=rows[0]; units=rows[1]
for v in rows[2:]:
    want=['Kernel Name','gpu__time_duration.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sectors.sum','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','smsp__cycles_active.avg','launch__grid_size','launch__block_size','launch__registers_per_thread','smsp__inst_executed.sum','smsp__issue_active.avg.per_cycle_active','l1tex__data_pipe_lsu_wavefronts.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__lsu_writeback_active.sum','l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum']
    for i,name in enumerate(h):
        if name in want or ('issue_stalled' in name and 'ratio' in name and float(v[i] or 0) > 0.5):
            print(f"{name:80s} {units[i]:10s} {v[i]}")
