#!/bin/bash
set -u
mkdir -p gpurun_out
stage() { name=$1; shift; echo "=== $name" ; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log; }
TAILN=15 stage fused_tests 600 python -m pytest tests/test_fused_gpu.py tests/test_kernels_gpu.py -q -x --no-header -p no:cacheprovider
TAILN=20 stage bench_fused 600 python tools/bench_fused.py
TAILN=20 stage bench_fused_noflush 600 python tools/bench_fused.py --no-flush
