#!/bin/bash
set -u
mkdir -p gpurun_out
stage() { name=$1; shift; echo "=== $name" ; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log; }
TAILN=30 stage loss_tracking 1200 python -m pytest tests/test_loss_tracking_gpu.py -q -x -s --no-header -p no:cacheprovider
