"""Summarise an .ncu-rep (ncu --set full) into the per-kernel CSV kept under profiles/.
usage: ncu_summary.py report.ncu-rep out.csv"""
import csv, subprocess, sys
COLS = """gpu__time_duration.sum launch__registers_per_thread launch__grid_size launch__block_size
launch__occupancy_limit_registers launch__occupancy_limit_shared_mem sm__warps_active.avg.pct_of_peak_sustained_active
smsp__inst_executed.sum sm__inst_executed.avg.per_cycle_elapsed smsp__issue_active.avg.pct_of_peak_sustained_active
smsp__thread_inst_executed_per_inst_executed.ratio sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
dram__bytes_read.sum dram__bytes_write.sum l1tex__t_sector_hit_rate.pct lts__t_sector_hit_rate.pct
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio""".split()
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units, data = rows[0], rows[1], rows[2:]
ix = {n: i for i, n in enumerate(h)}
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["Kernel Name"] + COLS)
    for r in data:
        w.writerow([r[ix["Kernel Name"]]] + [("%s %s" % (r[ix[c]], units[ix[c]])).strip() if c in ix else "" for c in COLS])
