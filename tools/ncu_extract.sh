#!/bin/bash
# usage: ncu_extract.sh <report.ncu-rep> <out.csv>: keeps the handful of raw metrics the docs cite (run on the GPU box, then
# delete the report: gpurun only brings 64 MiB back)
ncu -i "$1" --page raw --csv 2>/dev/null | python3 -c '
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr,units,vals=rows[0],rows[1],rows[2]
keep=["Kernel Name","gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed","sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active","sm__issue_active.avg.pct_of_peak_sustained_elapsed","sm__inst_executed.avg.pct_of_peak_sustained_elapsed","l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed","l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed","lts__t_sector_hit_rate.pct","launch__registers_per_thread","launch__grid_size","launch__block_size","launch__cluster_size","sm__warps_active.avg.pct_of_peak_sustained_active","l1tex__m_xbar2l1tex_read_bytes.sum"]
w=csv.writer(sys.stdout)
for h,u,v in zip(hdr,units,vals):
    if h in keep: w.writerow([h,u,v])
' > "$2"
