#!/bin/bash
set -u
OUT=gpurun_out/i6; mkdir -p $OUT
timeout 400 python -m pytest tests -m gpu -q --maxfail=12 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -6 $OUT/pytest_gpu.log
B="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
for w in c4 c5; do timeout 200 python bench.py --workload $w $B > $OUT/bench_$w.json 2> $OUT/bench_$w.err; echo "$w rc=$? $(cut -c1-130 $OUT/bench_$w.json)"; done
NDS_LIB_PATH=$PWD/ndarray-stats_b200/libndstats_b200_strvar.so timeout 200 python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/diag.log 2>&1; grep -c strtime $OUT/diag.log
SEEDS=1 timeout 200 bash tools/stress_bracket.sh > $OUT/stress.log 2>&1; grep -c "^==" $OUT/stress.log; grep -E "^bad|rror" $OUT/stress.log | sort | uniq -c | head
