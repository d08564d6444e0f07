"""usage: ncu -i rep --page raw --csv | python tools/ncu_rows.py > out.txt : one block per distinct kernel (its LAST
captured launch): the handful of metrics that say what an HBM-bound helper is waiting for."""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "smsp__inst_executed.sum", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]
last = {}
for r in rows[2:]:
    last[r[col["Kernel Name"]]] = r
for name, r in last.items():
    print("==", name)
    for k in KEEP:
        if k in col:
            print(f"   {k:75s} {r[col[k]]} {rows[1][col[k]]}")
    stalls = [(float(r[i].replace(",", "")), h) for h, i in col.items()
              if "issue_stalled" in h and h.endswith("per_warp_active.pct") and r[i] not in ("", "n/a")]
    for v, h in sorted(stalls, reverse=True)[:5]:
        print(f"   stall {h.split('issue_stalled_')[1].split('_per_warp')[0]:40s} {v:.1f} % of warp-cycles")
