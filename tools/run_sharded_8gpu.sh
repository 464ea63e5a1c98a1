#!/bin/bash
# 8-GPU session: the C4 strong-scaling record over peer memory and over NCCL, stage timings.
# usage: gpurun --gpus 8 --timeout 600 -- 'bash tools/run_sharded_8gpu.sh'
set -u
N=${N:-8}
OUT=gpurun_out/sh$N; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29513 bench.py --gpus $N --workload sharded_1d --steps 10 --warmup 3 > $OUT/bench_p2p.json 2> $OUT/bench_p2p.err; echo "bench p2p rc=$?"; cut -c1-300 $OUT/bench_p2p.json
NDS_XCHG=nccl timeout 300 $TR --master-port 29514 bench.py --gpus $N --workload sharded_1d --steps 10 --warmup 3 --no-parity > $OUT/bench_nccl.json 2> $OUT/bench_nccl.err; echo "bench nccl rc=$?"; cut -c1-300 $OUT/bench_nccl.json
NDS_BRK_TIMING=1 timeout 300 $TR --master-port 29515 bench.py --gpus $N --workload sharded_1d --steps 3 --warmup 3 --no-parity > $OUT/bench_timing.json 2> $OUT/bench_timing.err; echo "timing rc=$?"; grep "bracket-timing" $OUT/bench_timing.err | grep "count+S0" | tail -3
if [ "${CHECK:-0}" = "1" ]; then
  timeout 300 $TR --master-port 29511 tools/check_sharded_nccl.py > $OUT/check_p2p.log 2>&1; echo "check p2p rc=$?"; tail -1 $OUT/check_p2p.log
fi
