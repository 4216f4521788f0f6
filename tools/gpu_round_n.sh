#!/bin/bash
set -u
mkdir -p gpurun_out
python tools/run_conv_once.py 32 fwd 32 64 3 256 > gpurun_out/plain_conv2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_tc_fwd -s 2 -c 1 -o gpurun_out/ncu_conv2_fwd \
    python tools/run_conv_once.py 32 fwd 32 64 3 256 > gpurun_out/ncu_conv2_fwd.log 2>&1
echo "rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 1 --steps 4 --warmup 3 --no-cpu-baseline --no-classifier > gpurun_out/bench_torchrun1.log 2>&1
echo "rc=$?"; tail -n 1 gpurun_out/bench_torchrun1.log | cut -c1-200
