#!/usr/bin/env python
"""Static instruction mix of the main loop (largest backward branch) of one kernel, from `nvdisasm --print-line-info` output.
usage: python tools/sass_loop_mix.py <nvdisasm.txt> <kernel-name-substring>"""
import bisect
import collections
import re
import sys

txt, pat = sys.argv[1], sys.argv[2]
lines, inside = [], False
for l in open(txt):
    if l.startswith(".text."):
        inside = pat in l
        continue
    if inside:
        lines.append(l.rstrip("\n"))
labels = {m.group(1): i for i, l in enumerate(lines) for m in [re.match(r"^(\.L_x_\d+):", l)] if m}
best = None
for i, l in enumerate(lines):
    m = re.search(r"\bBRA(?:\.U)?\s+(?:!?U?P\w+,\s*)?`\((\.L_x_\d+)\)", l)
    if m and m.group(1) in labels and labels[m.group(1)] < i:
        span = i - labels[m.group(1)]
        if best is None or span > best[0]:
            best = (span, labels[m.group(1)], i)
_, start, end = best
cur = None
for i in range(start, -1, -1):
    m = re.search(r'//## File "([^"]+)", line (\d+)', lines[i])
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); break
ops, byfile, n = collections.Counter(), collections.Counter(), 0
for l in lines[start:end + 1]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\w+ )?(\S+)", l)
    if m:
        n += 1; ops[m.group(2).split(".")[0]] += 1; byfile[cur[0] if cur else "?"] += 1
print("kernel instructions: %d   main loop: %d (%.1f KB)" % (sum(1 for l in lines if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l)), n, n * 16 / 1024))
print("ops:", ", ".join("%s %d" % kv for kv in ops.most_common(16)))
print("files:", ", ".join("%s %d" % kv for kv in byfile.most_common()))
