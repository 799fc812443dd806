#!/usr/bin/env python3
"""Condenses an .ncu-rep (ncu -i ... --page raw --csv) into the handful of counters DESIGN.md cites.
usage: tools/ncu_summary.py report.ncu-rep [> profiles/<name>.txt]"""
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread ", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum ", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum ", "dram__bytes_write.sum ", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum ", "l1tex__t_bytes.sum ",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum", "smsp__inst_executed_op_global_ld.sum", "smsp__inst_executed_op_shared_ld.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg ", "sm__cycles_elapsed.avg ", "smsp__warp_issue_stalled", "smsp__average_warp", "smsp__warps_issue_stalled",
        "sass__inst_executed_global_loads", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "launch__local", "derived__", "sm__sass_thread_inst_executed_op_integer"]
def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("== kernel:", d.get("Kernel Name"), "grid", d.get("Grid Size"), "block", d.get("Block Size"))
        for h, u, v in zip(hdr, units, r):
            if any((h + " ").startswith(k) or k in h for k in KEYS):
                print("  %-90s %-14s %s" % (h, u, v))
if __name__ == "__main__":
    main(sys.argv[1])
