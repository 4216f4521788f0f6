#!/bin/bash
set -u
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:dwf_bwd -s 3 -c 1 -o gpurun_out/ncu_dwf_bwd4 \
    python tools/bench_fused.py --only "affine+sums+side" --reps 2 > gpurun_out/ncu_dwf_bwd4.log 2>&1
echo "rc=$?"
ncu --set full --clock-control none --import-source on -k regex:dwf_fwd -s 3 -c 1 -o gpurun_out/ncu_dwf_fwd4 \
    python tools/bench_fused.py --only "dwf_fwd affine" --reps 2 > gpurun_out/ncu_dwf_fwd4.log 2>&1
echo "rc=$?"
