#!/bin/bash
set -u
mkdir -p gpurun_out
stage() { name=$1; shift; echo "=== $name" ; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log; }
TAILN=12 stage tc_tests 600 python -m pytest tests/test_tc_gpu.py -q -x --no-header -p no:cacheprovider
for kp in 1 2 4; do
CERVIX_TC_PAIRS=$kp TAILN=30 stage conv_shapes_p$kp 600 python tools/bench_conv_shapes.py
done
