#!/bin/bash
set -u
mkdir -p gpurun_out
stage() { name=$1; shift; echo "=== $name" ; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log; }
TAILN=30 stage r2n_split_test 600 python -m pytest tests/test_engine_gpu.py -m gpu -q --no-header -p no:cacheprovider -k "split or fit_one"
