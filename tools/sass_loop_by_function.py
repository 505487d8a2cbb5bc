#!/usr/bin/env python
"""Static instruction count of a kernel's main loop (largest backward branch) per source function, from `nvdisasm --print-line-info`.
usage: python tools/sass_loop_by_function.py <nvdisasm.txt> <kernel-name-substring>"""
import bisect
import collections
import os
import re
import sys

txt, pat = sys.argv[1], sys.argv[2]
CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "optimization_dynamics_b200", "csrc")
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


def funcs(path):
    out = []
    for n, l in enumerate(open(path), 1):
        m = re.search(r"OD_HD (?:static )?[\w<>:,&\* ]*?\b(\w+)\(", l)
        if m and not l.strip().startswith("//"):
            out.append((n, m.group(1)))
    return out


F = {f: funcs(os.path.join(CSRC, f)) for f in ("contact_ip.cuh", "group_gj.cuh", "fastmath.cuh")}
cnt, cur, ops = collections.Counter(), None, collections.Counter()
for l in lines[start:end + 1]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\w+ )?(\S+)", l)
    if m and cur:
        f, ln = cur
        name = f
        if f in F:
            idx = bisect.bisect_right([a for a, _ in F[f]], ln) - 1
            name = f + ":" + (F[f][idx][1] if idx >= 0 else "?")
        cnt[name] += 1; ops[m.group(2).split(".")[0]] += 1
tot = sum(cnt.values())
print("main loop: %d instructions (%.1f KB)" % (tot, tot * 16 / 1024))
for k, v in cnt.most_common(30):
    print("%6d  %4.1f%%  %s" % (v, 100.0 * v / tot, k))
print("ops:", ", ".join("%s %d" % kv for kv in ops.most_common(14)))
