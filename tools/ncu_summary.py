"""Summarise an .ncu-rep (raw page) into the metric list kept under profiles/."""
import csv, subprocess, sys, io
rep, out = sys.argv[1], sys.argv[2]
kfilter = sys.argv[3] if len(sys.argv) > 3 else ""
if rep.endswith(".csv"):   # a raw page exported on the GPU box (tools/ncu_r02.sh)
    raw = open(rep).read()
else:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "lts__t_bytes.sum",
        "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"]
idx = {h: i for i, h in enumerate(hdr)}
with open(out, "w") as f:
    f.write(f"# {rep.split('/')[-1]}: ncu --set full --clock-control none --import-source on (per launch)\n")
    n = 0
    for r in rows[2:]:
        if kfilter and kfilter not in r[idx["Kernel Name"]]:
            continue
        f.write(f"--- launch {n}\n"); n += 1
        for w in WANT:
            if w in idx:
                f.write(f"  {w} [{units[idx[w]]}] = {r[idx[w]]}\n")
print("wrote", out)
