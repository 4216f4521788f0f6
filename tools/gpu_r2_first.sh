#!/bin/bash
# round 2, first GPU call: whole GPU suite WITHOUT -x, smoke, run-to-run spread probe, racecheck of the small step, bench
set -u
mkdir -p gpurun_out
stage() { name=$1; shift; echo "=== $name" ; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log; }
TAILN=25 stage r2_gpu_tests_all 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --durations=15
TAILN=8 stage r2_smoke 600 python -c "import __graft_entry__ as g; g.smoke()"
TAILN=12 stage r2_determinism 900 python tools/determinism_probe.py
TAILN=12 stage r2_racecheck 600 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/determinism_probe.py --reps 1 --cases bf16:64:4
TAILN=1 stage r2_bench_default 1200 python bench.py
