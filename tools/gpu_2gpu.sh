#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/bench_2gpu.log 2>&1
echo "rc=$?"; tail -n 2 gpurun_out/bench_2gpu.log | cut -c1-600
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_2gpu.log 2>&1
echo "rc=$?"; tail -n 1 gpurun_out/bench_ref_2gpu.log | cut -c1-400
