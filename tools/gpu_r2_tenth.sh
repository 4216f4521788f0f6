#!/bin/bash
set -u
mkdir -p gpurun_out
stage() { name=$1; shift; echo "=== $name" ; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log; }
TAILN=12 stage r2k_fit_tests 900 python -m pytest tests/test_predict_gpu.py tests/test_engine_gpu.py -m gpu -q --no-header -p no:cacheprovider
TAILN=6 stage r2k_rest 900 python -m pytest tests/test_tc_gpu.py tests/test_loss_tracking_gpu.py tests/test_patch_encoder.py -m gpu -q -x --no-header -p no:cacheprovider
