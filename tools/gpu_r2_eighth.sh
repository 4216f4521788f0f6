#!/bin/bash
set -u
mkdir -p gpurun_out
stage() { name=$1; shift; echo "=== $name" ; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log; }
TAILN=6 stage r2i_tests 900 python -m pytest tests/test_fusion_gpu.py tests/test_fused_gpu.py -m gpu -q -x --no-header -p no:cacheprovider
TAILN=4 stage r2i_head_bm_auto 300 python tools/profile_head.py
CERVIX_GEMM_BM=64 TAILN=4 stage r2i_head_bm64 300 python tools/profile_head.py
TAILN=3 stage r2i_colred3 200 python tools/bench_fused.py --only bn_bwd_sums --graph
CERVIX_COLRED_MINB4=1 TAILN=3 stage r2i_colred4 200 python tools/bench_fused.py --only bn_bwd_sums --graph
TAILN=1 stage r2i_bench 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline
