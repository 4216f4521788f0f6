#!/bin/bash
set -u
mkdir -p gpurun_out
stage() { name=$1; shift; echo "=== $name" ; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log; }
TAILN=4 stage r2x_tests 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_engine_gpu.py -m gpu -x -q --no-header -p no:cacheprovider
B="python bench.py --no-gpu-baseline --no-fit --no-classifier --no-cpu-baseline --no-augment --steps 12 --warmup 3"
TAILN=1 stage r2x_bench_on 300 $B
CERVIX_DROPOUT_MASK=1 TAILN=1 stage r2x_bench_off 300 $B
TAILN=1 stage r2x_bench_on2 300 $B
python - <<'PY'
import json
for n in ("on","off","on2"):
    l=[x for x in open("gpurun_out/r2x_bench_%s.log"%n) if x.startswith("{")][-1]
    d=json.loads(l); print(n, d["value"], d["ms_per_step"])
PY
