#!/usr/bin/env python
"""Print the handful of ncu raw-page counters this repo's notes quote, per captured launch: python tools/ncu_keys.py report.ncu-rep"""
import csv, io, subprocess, sys
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct"]
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    print("----", r[idx["Kernel Name"]][:110])
    for k in KEYS:
        if k in idx:
            print(f"  {k:75s} {r[idx[k]][:20]:>20s} {units[idx[k]]}")
    st = sorted(((float(r[idx[h]].replace(',', '') or 0), h[34:-23]) for h in stalls), reverse=True)[:7]
    print("  stalls/issue:", ", ".join(f"{n}={v:.2f}" for v, n in st))
