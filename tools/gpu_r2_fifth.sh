#!/bin/bash
set -u
mkdir -p gpurun_out
stage() { name=$1; shift; echo "=== $name" ; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log; }
TAILN=25 stage r2f_gpu_tests 1500 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider
for v in 0 1 2; do
CERVIX_DWF_VARIANT=$v TAILN=7 stage r2f_dwf_micro_v$v 300 python tools/bench_fused.py --only dwf_ --graph
done
for v in 0 1; do
CERVIX_DWF_VARIANT=$v TAILN=1 stage r2f_bench_v$v 600 python bench.py --steps 10 --warmup 3 --no-classifier --no-cpu-baseline --no-gpu-baseline
done
TAILN=30 stage r2f_head 300 python tools/profile_head.py
