"""Condenses `ncu --page raw --csv` exports into the small per-kernel summaries committed under profiles/.
Usage: python tools/ncu_summary.py raw.csv [metric-substring ...] -> CSV on stdout (one row per captured launch)"""
import csv
import sys

KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
        "sm__inst_executed_pipe_fmaheavy.sum", "smsp__inst_executed_pipe_fmaheavy.sum", "smsp__inst_executed_pipe_alu.sum",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]

path = sys.argv[1]
extra = sys.argv[2:]
rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
hdr, units, data = rows[0], rows[1], rows[2:]
cols = [i for i, h in enumerate(hdr) if h in KEEP or any(e in h for e in extra)]
name_i = hdr.index("Kernel Name")
w = csv.writer(sys.stdout)
w.writerow(["kernel"] + [f"{hdr[i]} [{units[i]}]" for i in cols])
for r in data:
    w.writerow([r[name_i][:70]] + [r[i] for i in cols])
