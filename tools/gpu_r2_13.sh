#!/bin/bash
# final 1-GPU validation: whole GPU suite, smoke, default bench, reference arm
set -u
mkdir -p gpurun_out
stage() { name=$1; shift; echo "=== $name" ; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log; }
TAILN=6 stage r2p_gpu_suite 1500 python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider
TAILN=3 stage r2p_smoke 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
TAILN=2 stage r2p_bench 900 python bench.py
TAILN=2 stage r2p_bench_ref 600 python bench.py --impl reference --steps 2 --warmup 1
