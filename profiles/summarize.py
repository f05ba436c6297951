#!/usr/bin/env python
"""Summarise an ncu report (read on the CPU box) into profiles/: per-kernel duration, DRAM bytes, occupancy, stalls.
usage: python profiles/summarize.py gpurun_out/prof.ncu-rep profiles/r01_<tag>"""
import csv
import io
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_active.avg", "sm__cycles_elapsed.avg", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
summ, traffic = [], {}
for r in rows[2:]:
    d = {"kernel": r[hdr.index("Kernel Name")]}
    for w in want:
        if w in hdr:
            d[w] = r[hdr.index(w)] + " " + units[hdr.index(w)]
    summ.append(d)
    try:
        def mb(name):
            v, u = float(r[hdr.index(name)].replace(",", "")), units[hdr.index(name)].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
        key = "k1_kstrongest" if "k1_" in d["kernel"] else "k3_surface_points" if "k3_" in d["kernel"] else "k5_register" if "k5_" in d["kernel"] else d["kernel"]
        traffic[key] = mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum")
    except Exception:
        pass
json.dump({"report": rep, "kernels": summ, "dram_bytes_per_launch": traffic}, open(out + "_ncu_summary.json", "w"), indent=1)
print(json.dumps(traffic))
