#!/bin/bash
set -u
mkdir -p gpurun_out
stage() { name=$1; shift; echo "=== $name" ; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log; }
TAILN=3 stage fused_tests 900 python -m pytest tests/test_fused_gpu.py tests/test_model_gpu.py -q -x --no-header -p no:cacheprovider
TAILN=8 stage bench_fused_graph 600 python tools/bench_fused.py --graph --only dwf
TAILN=1 stage bench_b32 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-classifier
