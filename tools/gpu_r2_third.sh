#!/bin/bash
set -u
mkdir -p gpurun_out
stage() { name=$1; shift; echo "=== $name" ; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log; }
TAILN=40 stage r2c_engine_tests 900 python -m pytest tests/test_engine_gpu.py tests/test_model_gpu.py tests/test_predict_gpu.py -q -x --no-header -p no:cacheprovider
TAILN=3 stage r2c_smoke 600 python -c "import __graft_entry__ as g; g.smoke()"
TAILN=1 stage r2c_bench 1500 python bench.py
