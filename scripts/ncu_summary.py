"""Summarise an .ncu-rep (first kernel): key throughput metrics + stall mix + hottest source lines.
usage: python scripts/ncu_summary.py file.ncu-rep [n_lines]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
h, v = r[0], r[2]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread",
        "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "smsp__average_warp_latency_per_inst_issued.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for i, k in enumerate(h):
    if k in want or (k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio")
                     and "not_issued" not in k and float(v[i] or 0) > 0.2):
        print(f"{k:90s} {v[i]} {r[1][i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, x in enumerate(rows) if x and x[0] == "Line No")
hd = rows[hi]
isamp, iex = hd.index("# Samples"), hd.index("Instructions Executed")
body = [x for x in rows[hi + 1:] if len(x) > isamp and x[0].isdigit() and x[isamp].isdigit()]
tot = sum(int(x[isamp]) for x in body) or 1
toti = sum(int(x[iex]) for x in body if x[iex].isdigit()) or 1
print(f"# hottest source lines (of {tot} samples, {toti} warp instructions)")
for x in sorted(body, key=lambda x: -int(x[isamp]))[:nl]:
    print(f"{100 * int(x[isamp]) / tot:5.1f}%  inst={100 * int(x[iex]) / toti:5.1f}%  L{x[0]:>4s}  {x[1].strip()[:110]}")
