#!/bin/bash
set -u
mkdir -p gpurun_out
stage() { name=$1; shift; echo "=== $name" ; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log; }
TAILN=12 stage model_tests 900 python -m pytest tests/test_model_gpu.py -q -x --no-header -p no:cacheprovider
TAILN=1 stage bench_b32 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-classifier
