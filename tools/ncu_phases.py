#!/usr/bin/env python
"""Stall samples and executed warp instructions of one kernel grouped by ranges of CUDA source lines of ONE file
(SASS walked in address order; helper lines inlined from other places count for the range they sit in).
usage: tools/ncu_phases.py X.ncu-rep lo:hi:name [lo:hi:name ...]"""
import csv, io, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]
phases = [(int(a.split(":")[0]), int(a.split(":")[1]), a.split(":")[2]) for a in sys.argv[2:]]
LO, HI = min(p[0] for p in phases), max(p[1] for p in phases)
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
for i, r in enumerate(rows):
    if "# Samples" in r:
        hdr, start = r, i + 1
        break
S, EX, A = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Address")
cur, sass = None, []
for r in rows[start:]:
    if len(r) <= EX: continue
    if r[0].strip():
        try: cur = int(r[0])
        except ValueError: pass
        continue
    try: sass.append((int(r[A], 16), int(r[S]), int(r[EX]), cur))
    except ValueError: pass
sass.sort()
dedup, seen = [], set()   # an inlined instruction is listed under every CUDA line of its inline stack: keep one copy,
for x in sass:            # tagged with the line of the kernel's own file when there is one
    if x[0] in seen:
        if x[3] is not None and LO <= x[3] <= HI and not (dedup[-1][3] is not None and LO <= dedup[-1][3] <= HI): dedup[-1] = x
        continue
    seen.add(x[0]); dedup.append(x)
sass = dedup
tot = sum(x[1] for x in sass) or 1
agg, last = defaultdict(lambda: [0, 0]), 0
for a, s, e, l in sass:
    major = l if (l is not None and LO <= l <= HI) else last
    last = major
    for lo, hi, name in phases:
        if lo <= major <= hi:
            agg[name][0] += s; agg[name][1] += e
            break
    else:
        agg["other"][0] += s; agg["other"][1] += e
for name in [p[2] for p in phases] + ["other"]:
    v = agg[name]
    print(f"{name:24s} samples {100 * v[0] / tot:5.1f}%   warp-instr {v[1] / 1e6:8.1f}M")
