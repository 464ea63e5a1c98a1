#!/usr/bin/env python
"""SASS of one kernel inside a range of CUDA source lines (address order), with samples and the main stall reason.
usage: tools/ncu_range.py X.ncu-rep LO HI [min_samples]"""
import csv, io, subprocess, sys
rep, LO, HI = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
mins = int(sys.argv[4]) if len(sys.argv) > 4 else 0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
for i, r in enumerate(rows):
    if "# Samples" in r:
        hdr, start = r, i + 1
        break
S, EX, A = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Address")
stall = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
cur, sass = None, {}
for r in rows[start:]:
    if len(r) <= EX: continue
    if r[0].strip():
        try: cur = int(r[0])
        except ValueError: pass
        continue
    try: a = int(r[A], 16)
    except ValueError: continue
    own = cur is not None and LO <= cur <= HI
    if a not in sass or (own and not sass[a][0]): sass[a] = (own, cur, r)
tot = 0
for a in sorted(sass):
    own, l, r = sass[a]
    if not own: continue
    s = int(r[S] or 0); tot += s
    if s < mins: continue
    top = max(stall, key=lambda ih: int(r[ih[0]] or 0))
    print(f"{s:6d} exec {int(r[EX] or 0):>9d} L{l:<4d} {top[1][6:]:12s} {r[3].strip()[:90]}")
print("samples in range:", tot)
