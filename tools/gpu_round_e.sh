#!/bin/bash
set -u
mkdir -p gpurun_out
stage() { name=$1; shift; echo "=== $name" ; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log; }
TAILN=8 stage fused_tests 900 python -m pytest tests/test_fused_gpu.py tests/test_tc_gpu.py tests/test_kernels_gpu.py tests/test_model_gpu.py -q -x --no-header -p no:cacheprovider
TAILN=6 stage bench_fused 600 python tools/bench_fused.py --only conv_fwd
CERVIX_STATS_EPILOGUE=0 TAILN=1 stage bench_b32_nostats 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline
TAILN=1 stage bench_b32 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline
