import sys,csv
want=("gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","lts__t_bytes.sum","lts__throughput.avg.pct_of_peak_sustained_elapsed","l1tex__throughput.avg.pct_of_peak_sustained_elapsed","lts__t_sector_hit_rate.pct","sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active","smsp__issue_active.avg.pct_of_peak_sustained_active","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","launch__registers_per_thread","gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed","sm__throughput.avg.pct_of_peak_sustained_elapsed","l1tex__lsu_writeback_active_mem_lg.sum","lts__t_sectors_srcunit_tex.sum","lts__t_sectors_op_read.sum","sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active","smsp__inst_executed.sum","l1tex__m_xbar2l1tex_read_bytes.sum","l1tex__m_xbar2l1tex_read_bytes.sum.per_second","lts__t_sectors.avg.pct_of_peak_sustained_elapsed","lts__d_sectors.avg.pct_of_peak_sustained_elapsed","lts__xbar2lts_cycles_active.avg.pct_of_peak_sustained_elapsed","lts__lts2xbar_cycles_active.avg.pct_of_peak_sustained_elapsed")
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[0]; 
for r in rows[2:]:
    print("kernel:", r[hdr.index("Kernel Name")][:60])
    for i,h in enumerate(hdr):
        if h in want or any(k in h for k in ("lts__throughput","xbar","l1tex__throughput")):
            print(f"  {h:80s} {r[i]} {rows[1][i]}")
