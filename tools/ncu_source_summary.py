#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv` of one kernel: executed warp instructions by opcode, the
hottest SASS lines by stall samples, lane efficiency.  usage: tools/ncu_source_summary.py X.ncu-rep [topN]"""
import csv, io, subprocess, sys
from collections import Counter
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
S, EX, TEX, SMP = ci["Source"], ci["Instructions Executed"], ci["Thread Instructions Executed"], ci["# Samples"]
tot = tex = smp = 0; ops = Counter(); lines = []
for n, r in enumerate(rows[2:]):
    try: v = int(r[EX]); t = int(r[TEX]); s = int(r[SMP])
    except (ValueError, IndexError): continue
    tot += v; tex += t; smp += s
    tk = r[S].split(); op = tk[0] if tk else "?"
    if op.startswith("@") and len(tk) > 1: op = tk[1]
    ops[op.split(".")[0]] += v
    lines.append((s, v, n, r[S].strip()[:100]))
print(f"warp instructions {tot}  thread instructions {tex}  lanes/instr {tex / max(tot, 1):.1f}  samples {smp}")
for op, v in ops.most_common(topn): print(f"  {op:10s} {v:>12d} {100.0 * v / tot:5.1f}%")
print("hottest lines by stall samples:")
for s, v, n, src in sorted(lines, reverse=True)[:topn]: print(f"  {100.0 * s / max(smp, 1):5.1f}% samples  exec {v:>10d}  #{n:<5d} {src}")
