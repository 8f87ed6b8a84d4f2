#!/usr/bin/env python3
"""Short summary of an .ncu-rep (run here, no GPU needed):  python tools/ncu_summary.py file.ncu-rep [more...]"""
import csv, subprocess, sys
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__cycles_elapsed.avg.per_second", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "sm__icc_request_hit_rate.pct", "idc__request_hit_rate.pct"]
for f in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", f, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print(f"== {f}: {vals[hdr.index('Kernel Name')]}")
        for i, h in enumerate(hdr):
            if h in WANT or ("issue_stalled" in h and h.endswith("per_issue_active.ratio") and float(vals[i] or 0) > 0.02):
                print(f"  {h} [{units[i]}] = {vals[i]}")
