#!/bin/bash
set -u
mkdir -p gpurun_out
stage() { name=$1; shift; echo "=== $name" ; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log; }
TAILN=60 stage r2d_new_tests 1500 python -m pytest tests/test_engine_gpu.py tests/test_baseline_configs_gpu.py tests/test_patch_encoder.py -q -s --no-header -p no:cacheprovider -m gpu --durations=8
