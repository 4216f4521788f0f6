#!/bin/bash
# 2-GPU: the default data-parallel bench line with the strong-scaling leg forced at 64 images per GPU
set -u
mkdir -p gpurun_out
stage() { name=$1; shift; echo "=== $name" ; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621"
CERVIX_BENCH_STRONG_BSZ=64 TAILN=1 stage r2s_bench2 900 $TR bench.py --gpus 2 --steps 8 --warmup 3
python - <<'PY'
import json
l=[x for x in open("gpurun_out/r2s_bench2.log") if x.startswith("{")][-1]
d=json.loads(l)
print("value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["step_gaps"], "f32", d["e2e"]["fp32_contract"]["value"])
print("strong", d["configs3_global_batch_256"])
print("classifier", {k: d["classifier"][k] for k in ("patients_per_s", "ms_per_step")} if d.get("classifier") else None)
print("allreduce", d["config"]["allreduce"][:60])
PY
