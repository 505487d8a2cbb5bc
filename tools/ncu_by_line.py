#!/usr/bin/env python
"""Aggregate an ncu report's SASS-level samples / executed instructions by CUDA source line.
usage: python tools/ncu_by_line.py gpurun_out/prof.ncu-rep [top_n]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = None
agg = {}
fname = "?"
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if "# Samples" in r and "Instructions Executed" in r:
        hdr = r; isamp = r.index("# Samples"); iexe = r.index("Instructions Executed"); continue
    if hdr is None or len(r) <= max(isamp, iexe):
        continue
    if r[0].isdigit():          # a CUDA source line with its aggregated metrics
        try:
            key = (fname, int(r[0]), r[1].strip()[:110])
            v = agg.setdefault(key, [0, 0])
            v[0] += int(r[isamp] or 0); v[1] += int(r[iexe] or 0)
        except ValueError:
            pass
tot_s = sum(v[0] for v in agg.values()); tot_e = sum(v[1] for v in agg.values())
print("total samples %d, executed warp-instr %d" % (tot_s, tot_e))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% samp %5.1f%% exec | %s:%d  %s" % (100.0 * v[0] / max(tot_s, 1), 100.0 * v[1] / max(tot_e, 1), k[0], k[1], k[2]))
