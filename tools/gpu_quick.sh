#!/bin/bash
set -u
mkdir -p gpurun_out
stage() { name=$1; shift; echo "=== $name" ; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log; }
stage kernels 600 python -m pytest tests/test_kernels_gpu.py tests/test_tc_gpu.py tests/test_model_gpu.py -q --no-header -p no:cacheprovider
TAILN=10 stage bench_kernels 600 python tools/bench_kernels.py
stage bench_b32_nograph 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-graph
stage bench_b32 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline
