#!/usr/bin/env python
"""Hot CUDA source lines of one kernel in an ncu report (needs -lineinfo and --import-source on):
   tools/ncu_lines.py X.ncu-rep <kernel regex> [topN] [launch-skip]"""
import csv, io, subprocess, sys
from collections import defaultdict
rep, pat = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
skip = sys.argv[4] if len(sys.argv) > 4 else "0"
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      f"regex:{pat}", "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
files = {}
cur_file = None
hdr = None
agg = defaultdict(lambda: [0, 0, ""])
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        smp, ex = hdr.index("# Samples"), hdr.index("Instructions Executed")
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    try:
        line = int(r[0])
        s, e = int(r[smp]), int(r[ex])
    except ValueError:
        continue
    k = (cur_file, line)
    agg[k][0] += s
    agg[k][1] += e
    if r[1] and not agg[k][2]:
        agg[k][2] = r[1].strip()[:110]
tot = sum(v[0] for v in agg.values()) or 1
tex = sum(v[1] for v in agg.values())
print(f"samples {tot}  warp instructions {tex}")
for (f, line), (s, e, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{100.0 * s / tot:5.1f}%  exec {e:>9d}  {f}:{line:<5d} {src}")
