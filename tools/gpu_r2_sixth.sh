#!/bin/bash
set -u
mkdir -p gpurun_out
stage() { name=$1; shift; echo "=== $name" ; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log; }
TAILN=12 stage r2g_tests_pdl 900 python -m pytest tests/test_fused_gpu.py tests/test_kernels_gpu.py tests/test_tc_gpu.py tests/test_model_gpu.py tests/test_fusion_gpu.py tests/test_engine_gpu.py -m gpu -q -x --no-header -p no:cacheprovider
for v in 1 3 4; do
CERVIX_DWF_VARIANT=$v TAILN=6 stage r2g_dwf_micro_v$v 300 python tools/bench_fused.py --only dwf_ --graph
done
CERVIX_PDL=1 TAILN=1 stage r2g_bench_pdl1 600 python bench.py --steps 10 --warmup 3 --no-classifier --no-cpu-baseline --no-gpu-baseline
CERVIX_PDL=0 TAILN=1 stage r2g_bench_pdl0 600 python bench.py --steps 10 --warmup 3 --no-classifier --no-cpu-baseline --no-gpu-baseline
TAILN=12 stage r2g_head 300 python tools/profile_head.py
