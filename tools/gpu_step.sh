#!/bin/bash
# one quick confirmation pass: the tests named in $TESTS (default: the patch-encoder file), then the default bench line
set -u
mkdir -p gpurun_out
stage() { name=$1; shift; echo "=== $name" ; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log; }
stage step_tests 600 python -m pytest ${TESTS:-tests/test_patch_encoder.py} -q --no-header -p no:cacheprovider -m gpu
TAILN=1 stage step_bench 900 python bench.py ${BENCH_ARGS:---no-cpu-baseline}
