#!/bin/bash
set -u
mkdir -p gpurun_out
stage() { name=$1; shift; echo "=== $name" ; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log; }
TAILN=4 stage gpu_tests_all 1800 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider
TAILN=3 stage smoke 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
TAILN=1 stage bench_default 1200 python bench.py
TAILN=1 stage bench_ref 600 python bench.py --impl reference --steps 2 --warmup 1
