#!/usr/bin/env python
"""Summarise an ncu report (one kernel launch) into profiles/: selected raw metrics (JSON) + hot source lines (text).
usage: python tools/ncu_summary.py gpurun_out/<rep>.ncu-rep profiles/<name>"""
import csv
import io
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keep = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__waves_per_multiprocessor", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sass__inst_executed_shared_loads",
        "sass__inst_executed_shared_stores", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "sm__cycles_active.avg",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
res = []
for r in rows[2:]:
    d = {}
    for h, u, v in zip(hdr, units, r):
        if h in keep:
            d[h + (" [%s]" % u if u else "")] = v
    res.append(d)


def num(d, key):
    for k, v in d.items():
        if k.startswith(key):
            try:
                x = float(v.replace(",", ""))
            except ValueError:
                return None
            if "[Kbyte]" in k: x *= 1e3
            if "[Mbyte]" in k: x *= 1e6
            if "[Gbyte]" in k: x *= 1e9
            return x
    return None


summ = {"report": rep, "launches": res}
if res:
    rd, wr = num(res[0], "dram__bytes_read.sum"), num(res[0], "dram__bytes_write.sum")
    summ["dram_bytes_per_launch"] = (rd or 0) + (wr or 0)
json.dump(summ, open(out + ".json", "w"), indent=1)
lines = subprocess.run([sys.executable, "tools/ncu_by_line.py", rep, "40"], stdout=subprocess.PIPE, text=True).stdout
open(out + "_hot_lines.txt", "w").write(lines)
print(json.dumps(summ, indent=1)[:3000])
