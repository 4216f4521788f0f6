#!/bin/bash
set -u
mkdir -p gpurun_out
stage() { name=$1; shift; echo "=== $name" ; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log; }
TAILN=6 stage r2t_tests 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_baseline_configs_gpu.py tests/test_engine_gpu.py -m gpu -x -q --no-header -p no:cacheprovider
TAILN=4 stage r2t_families 300 python tools/run_families_once.py --only upsample
TAILN=1 stage r2t_bench 600 python bench.py --no-gpu-baseline --no-fit --no-classifier --no-cpu-baseline --no-augment --steps 8 --warmup 3
python - <<'PY'
import json
l=[x for x in open("gpurun_out/r2t_bench.log") if x.startswith("{")][-1]
d=json.loads(l)
print("value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"])
PY
