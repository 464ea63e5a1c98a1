#!/bin/bash
# Round-2 evidence on one B200, most important first: GPU parity tests, smoke, the default bench (both arms), ncu launch
# lists and one `--set full` capture of each workload's dominant kernel.  usage: tools/evidence_r02.sh <tag> [tests|bench|launches|full ...]
# (gpurun brings back at most 64 MiB: run the stages in separate calls if the reports grow)
set -u
TAG=${1:-r02e}; OUT=gpurun_out/$TAG; mkdir -p $OUT
shift; STAGES="${*:-tests bench launches full}"
want() { case " $STAGES " in *" $1 "*) return 0;; *) return 1;; esac; }
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/gpu.csv 2>&1
if want tests; then
  timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log
  timeout 120 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log; tail -1 $OUT/smoke.log
fi
if want bench; then
  timeout 600 python bench.py > $OUT/bench_c2.json 2> $OUT/bench_c2.err; echo "bench default rc=$?"; cut -c1-200 $OUT/bench_c2.json
  timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_c2_reference.json 2>> $OUT/bench_c2.err; echo "ref rc=$?"
fi
want launches && for w in c2 c3 c4 c5 c5_i64_lanes_lower c1 x_mid_f64_rows x_tall_f32_cols; do
  CMD="python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity"
  timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$w.csv $CMD > $OUT/ncu_launches_$w.log 2>&1; echo "launches $w rc=$?"
done
want full && for wk in "c2 short_bitslice" "c3 cols_tma" "c4 brk_stream" "c5 brk_stream" "c5_i64_lanes_lower short_bitslice" "c1 short_bitslice" "x_mid_f64_rows mid_bitslice" "x_tall_f32_cols mid_bitslice"; do
  set -- $wk
  CMD="python bench.py --workload $1 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity"
  timeout 200 ncu --set full --clock-control none -k regex:$2 -s 3 -c 1 -f -o $OUT/prof_$1 $CMD > $OUT/ncu_full_$1.log 2>&1; echo "full $1 rc=$?"
  # keep the raw metric table, drop the report (gpurun brings back at most 64 MiB)
  ncu -i $OUT/prof_$1.ncu-rep --page raw --csv > $OUT/prof_$1.raw.csv 2>/dev/null && rm -f $OUT/prof_$1.ncu-rep
done
