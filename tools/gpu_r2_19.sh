#!/bin/bash
# gradient chains: parity tests, then A/B of the captured step on one box
set -u
mkdir -p gpurun_out
stage() { name=$1; shift; echo "=== $name" ; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log; }
TAILN=6 stage r2v_tests 900 python -m pytest tests/test_model_gpu.py tests/test_fused_gpu.py tests/test_baseline_configs_gpu.py -m gpu -x -q --no-header -p no:cacheprovider
B="python bench.py --no-gpu-baseline --no-fit --no-classifier --no-cpu-baseline --no-augment --steps 12 --warmup 3"
TAILN=1 stage r2v_bench_on 300 $B
CERVIX_GRAD_CHAIN=0 TAILN=1 stage r2v_bench_off 300 $B
CERVIX_GRAD_CHAIN=2 TAILN=1 stage r2v_bench_on2 300 $B
CERVIX_GRAD_CHAIN=0 TAILN=1 stage r2v_bench_off2 300 $B
python - <<'PY'
import json
for n in ("on","off","on2","off2"):
    l=[x for x in open("gpurun_out/r2v_bench_%s.log"%n) if x.startswith("{")][-1]
    d=json.loads(l); print(n, d["value"], d["ms_per_step"])
PY
CERVIX_GRAD_CHAIN=2 TAILN=1 stage r2v_prof_on 300 python tools/profile_step.py
grep -i "sum of kernel\|CUDAFunctor_add\|spatial_broadcast\|upsample_fwd" gpurun_out/r2v_prof_on.log | cut -c1-150
