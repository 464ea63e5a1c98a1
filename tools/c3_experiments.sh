#!/bin/bash
# C3 access-pattern experiments (1 GPU): dry streaming with boxes of 8 / 32 / 256 lanes, then the real kernels.
B="--workload c3 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-parity"
run() { python bench.py $B 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$1', round(d['ms_per_step'],4), round(d['roofline']['frac'],3))"; }
for w in 8 32 256; do NDS_COLS_TMA=2 NDS_COLS_DRY=$w run "dry box_w=$w slots=10"; done
for s in 4 6 16; do NDS_COLS_TMA=2 NDS_COLS_DRY=8 NDS_COLS_SLOTS=$s run "dry box_w=8 slots=$s"; done
for s in 6 8 10; do NDS_COLS_TMA=2 NDS_COLS_SLOTS=$s run "tma slots=$s"; done
NDS_COLS_TMA=1 run "ldgsts producer"
