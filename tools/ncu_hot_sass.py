#!/usr/bin/env python
"""SASS instructions of one kernel executed at least MIN times (address order), with the CUDA line they come from, their
samples and executed count: the hot loop as the hardware ran it.  usage: tools/ncu_hot_sass.py X.ncu-rep MIN_EXEC"""
import csv, io, subprocess, sys
rep, mine = sys.argv[1], int(float(sys.argv[2]))
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
for i, r in enumerate(rows):
    if "# Samples" in r:
        hdr, start = r, i + 1
        break
S, EX, A, SRC = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Address"), hdr.index("Address") + 1
cur, sass = None, {}
for r in rows[start:]:
    if len(r) <= EX: continue
    if r[0].strip():
        try: cur = int(r[0])
        except ValueError: pass
        continue
    try: a = int(r[A], 16)
    except ValueError: continue
    if a not in sass: sass[a] = (cur, r)
n = tot = 0
for a in sorted(sass):
    l, r = sass[a]
    e = int(r[EX] or 0)
    if e < mine: continue
    n += 1; tot += e
    print(f"{a & 0xfffff:05x} L{l or 0:<5d} s{int(r[S] or 0):>6d} e{e / 1e6:8.1f}M  {r[SRC].strip()[:90]}")
print(f"# {n} instructions, {tot / 1e6:.1f}M executed")
