#!/usr/bin/env python
"""Key metrics of the first kernel in an ncu report (read here, on the CPU box): usage tools/ncu_key.py report.ncu-rep [more reports...]"""
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__warps_eligible.avg.per_cycle_active"]
STALL = "smsp__average_warps_issue_stalled_"
for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"== {path} :: {d.get('Kernel Name', '?')}")
        for k in KEYS:
            if k in d:
                print(f"{k:90s} {d[k]:>18s} {units[hdr.index(k)]}")
        for h in hdr:
            if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and d[h] not in ("", "0"):
                try:
                    if float(d[h]) >= 0.05:
                        print(f"{h:90s} {float(d[h]):18.3f}")
                except ValueError:
                    pass
