#!/bin/bash
# usage: bash tools/gpu_round_ngpu.sh N
set -u
N=$1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/bench_${N}gpu.log 2>&1
echo "rc=$?"; tail -n 1 gpurun_out/bench_${N}gpu.log | cut -c1-260
