#!/usr/bin/env python
"""Static SASS instruction count per source line / section for one kernel (from nvdisasm --print-line-info of the cubin).
usage: python tools/sass_by_line.py <nvdisasm-line-info.txt> <kernel-name-substring> [top_n]"""
import collections
import re
import sys

txt, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
cnt = collections.Counter(); total = 0
inside = False; cur = ("?", 0)
for l in open(txt):
    if l.startswith(".text."):
        inside = pat in l
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        cnt[cur] += 1; total += 1
print("total SASS instructions:", total)
byfile = collections.Counter()
for (f, ln), c in cnt.items():
    byfile[f] += c
for f, c in byfile.most_common():
    print("%6d  %s" % (c, f))
print("--- top lines")
for (f, ln), c in cnt.most_common(top):
    print("%6d  %s:%d" % (c, f, ln))
