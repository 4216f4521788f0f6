#!/bin/bash
# end-of-round evidence: bench (default flags), graph-timed chain micro-benchmarks, per-shape conv table, step breakdown,
# ncu captures of the depthwise kernels
set -u
mkdir -p gpurun_out
stage() { name=$1; shift; echo "=== $name" ; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-3}" gpurun_out/$name.log; }
TAILN=1 stage bench_default 1200 python bench.py
TAILN=2 stage bench_fused_graph 600 python tools/bench_fused.py --graph
TAILN=2 stage conv_shapes 900 python tools/bench_conv_shapes.py
TAILN=2 stage prof_b32 900 python tools/profile_step.py
bash tools/gpu_ncu_dwf.sh
