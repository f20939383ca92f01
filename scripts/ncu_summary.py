#!/usr/bin/env python
"""Write a text summary of an .ncu-rep (key metrics + hottest source lines) for profiles/.  usage: ncu_summary.py rep out.txt"""
import csv, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_fp32.sum"]
with open(out, "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on  (%s)\n" % rep.split("/")[-1])
    for h, u, v in zip(hdr, units, vals):
        if h in want:
            f.write("%-78s %-14s %s\n" % (h, u, v))
    f.write("\n# hottest CUDA source lines by warp-stall samples\n")
    f.write(subprocess.run([sys.executable, "scripts/ncu_hot.py", rep, "14"], capture_output=True, text=True).stdout)
    f.write("\n# top CUDA source lines by executed warp instructions\n")
    f.write(subprocess.run([sys.executable, "scripts/ncu_inst.py", rep, "14"], capture_output=True, text=True).stdout)
print("wrote", out)
