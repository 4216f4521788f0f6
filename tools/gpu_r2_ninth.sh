#!/bin/bash
set -u
mkdir -p gpurun_out
stage() { name=$1; shift; echo "=== $name" ; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log; }
TAILN=8 stage r2j_gpu_tests_all 1500 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider --durations=6
TAILN=4 stage r2j_smoke 600 python -c "import __graft_entry__ as g; g.smoke()"
TAILN=1 stage r2j_bench_default 1500 python bench.py
CERVIX_CONV_BN_MINK=0 TAILN=1 stage r2j_bench_convbn0 600 python bench.py --steps 10 --warmup 3 --no-classifier --no-cpu-baseline --no-gpu-baseline --no-fit
CERVIX_CONV_BN_MINK=1024 TAILN=1 stage r2j_bench_convbn1024 600 python bench.py --steps 10 --warmup 3 --no-classifier --no-cpu-baseline --no-gpu-baseline --no-fit
TAILN=1 stage r2j_bench_ref 600 python bench.py --impl reference --steps 4 --warmup 1
