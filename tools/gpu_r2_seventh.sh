#!/bin/bash
set -u
mkdir -p gpurun_out
stage() { name=$1; shift; echo "=== $name" ; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log; }
TAILN=6 stage r2h_tests 900 python -m pytest tests/test_fusion_gpu.py tests/test_engine_gpu.py tests/test_baseline_configs_gpu.py tests/test_patch_encoder.py -m gpu -q -x --no-header -p no:cacheprovider
TAILN=34 stage r2h_head 300 python tools/profile_head.py
