#!/bin/bash
set -u
mkdir -p gpurun_out
stage() { name=$1; shift; echo "=== $name" ; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log; }
TAILN=15 stage tc_tests 600 python -m pytest tests/test_tc_gpu.py tests/test_kernels_gpu.py tests/test_fused_gpu.py -q -x --no-header -p no:cacheprovider
TAILN=40 stage conv_shapes 900 python tools/bench_conv_shapes.py
