"""Selected metrics of every kernel in an ncu report: `ncu -i rep --page raw --csv | python scripts/ncu_summary.py`."""
import csv, sys
KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct"]
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("----")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:85s} {r[i][:110]} {units[i]}")
    for i, k in enumerate(hdr):
        if "issue_stalled" in k and k.endswith("_per_issue_active.ratio"):
            try:
                v = float(r[i])
            except ValueError:
                continue
            if v >= 0.2:
                print(f"{k:85s} {v:.3f}")
