"""`.ncu-rep` -> the small metric table committed under profiles/ (one column per captured kernel).
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.csv"""
import csv, io, subprocess, sys

METRICS = """launch__grid_size launch__block_size launch__registers_per_thread gpu__time_duration.sum dram__bytes_read.sum
dram__bytes_write.sum gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed sm__throughput.avg.pct_of_peak_sustained_elapsed
sm__warps_active.avg.pct_of_peak_sustained_active smsp__inst_executed.sum smsp__issue_active.avg.pct_of_peak_sustained_active
smsp__average_warp_latency_per_inst_issued.ratio sm__cycles_elapsed.max launch__occupancy_limit_shared_mem
launch__occupancy_limit_registers l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio
smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio""".split()

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {n: i for i, n in enumerate(hdr)}
w = csv.writer(sys.stdout)
w.writerow(["metric", "unit"] + [r[ix["Kernel Name"]][:40] for r in data])
w.writerow(["Kernel Name", ""] + [r[ix["Kernel Name"]][:60] for r in data])
for m in METRICS:
    if m in ix:
        w.writerow([m, units[ix[m]]] + [r[ix[m]] for r in data])
