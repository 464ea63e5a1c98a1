#!/bin/bash
# GPU iteration: load flavour / L2 prefetch knobs of the streaming pass (NDS_STR_MODE)
set -u
OUT=gpurun_out/i4; mkdir -p $OUT
B="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
for mode in 0 1 0x20 0x40 0x41 0x80; do
  NDS_STR_MODE=$mode timeout 200 python bench.py --workload c4 $B > $OUT/bench_c4_m$mode.json 2> $OUT/bench_c4_m$mode.err; echo "mode=$mode c4 rc=$? $(cut -c1-130 $OUT/bench_c4_m$mode.json)"
done
for mode in 0 0x40; do
  NDS_STR_MODE=$mode timeout 200 python bench.py --workload c5 $B > $OUT/bench_c5_m$mode.json 2> $OUT/bench_c5_m$mode.err; echo "mode=$mode c5 rc=$? $(cut -c1-130 $OUT/bench_c5_m$mode.json)"
done
timeout 200 python bench.py --workload c3 $B > $OUT/bench_c3.json 2> $OUT/bench_c3.err; echo "c3 rc=$? $(cut -c1-130 $OUT/bench_c3.json)"
CMD="python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
NDS_STR_MODE=0x40 timeout 200 ncu --metrics gpu__time_duration.sum,sm__cycles_active.min,sm__cycles_active.avg,sm__cycles_active.max,smsp__inst_executed.sum --clock-control none -k regex:brk_stream -c 10 --csv --log-file $OUT/stream_m40.csv $CMD > $OUT/ncu_m40.log 2>&1
NDS_STR_MODE=0 timeout 200 ncu --metrics gpu__time_duration.sum,sm__cycles_active.min,sm__cycles_active.avg,sm__cycles_active.max,smsp__inst_executed.sum --clock-control none -k regex:brk_stream -c 10 --csv --log-file $OUT/stream_m0.csv $CMD > $OUT/ncu_m0.log 2>&1
grep -h "brk_stream_kernel<double, 1>" $OUT/stream_m40.csv | tail -5 | cut -d, -f5,13-
grep -h "brk_stream_kernel<double, 1>" $OUT/stream_m0.csv | tail -5 | cut -d, -f5,13-
