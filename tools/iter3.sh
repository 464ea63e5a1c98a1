#!/bin/bash
# GPU iteration: tile-rotation knob of the streaming pass (NDS_STR_ROT), cols issue path, parity
set -u
OUT=gpurun_out/i3; mkdir -p $OUT
timeout 400 python -m pytest tests -m gpu -q --maxfail=12 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -8 $OUT/pytest_gpu.log
B="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
for rot in 0 1 37; do for w in c4 c5; do
  NDS_STR_ROT=$rot timeout 200 python bench.py --workload $w $B > $OUT/bench_${w}_rot$rot.json 2> $OUT/bench_${w}_rot$rot.err; echo "rot=$rot $w rc=$? $(cut -c1-130 $OUT/bench_${w}_rot$rot.json)"
done; done
timeout 200 python bench.py --workload c3 $B > $OUT/bench_c3.json 2> $OUT/bench_c3.err; echo "c3 rc=$? $(cut -c1-130 $OUT/bench_c3.json)"
CMD="python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:brk_stream -s 3 -c 1 -f -o $OUT/prof_c4 $CMD > $OUT/ncu_full_c4.log 2>&1
SEEDS=1 timeout 200 bash tools/stress_bracket.sh > $OUT/stress.log 2>&1; grep -c "^==" $OUT/stress.log; grep -E "^bad|rror" $OUT/stress.log | sort | uniq -c | head
