#!/bin/bash
set -u
mkdir -p gpurun_out
stage() { name=$1; shift; echo "=== $name" ; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log; }
TAILN=3 stage r2r_bench_aug 600 python bench.py --no-gpu-baseline --no-fit --no-classifier --no-cpu-baseline --steps 4 --warmup 3
python - <<'PY'
import json
l=[x for x in open("gpurun_out/r2r_bench_aug.log") if x.startswith("{")][-1]
d=json.loads(l)
print(json.dumps(d.get("augment"), indent=1))
print("value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["step_gaps"], "f32", d["e2e"]["fp32_contract"]["value"], d["e2e"]["fp32_contract"]["step_gaps"])
PY
