#!/bin/bash
# A/B on one box: gradient chain of ASPP on / off - kernel-time breakdown and the captured step
set -u
mkdir -p gpurun_out
stage() { name=$1; shift; echo "=== $name" ; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log; }
TAILN=1 stage r2u_prof_on 300 python tools/profile_step.py
CERVIX_GRAD_CHAIN=0 TAILN=1 stage r2u_prof_off 300 python tools/profile_step.py
for m in on off; do echo "--- chain $m"; grep -i "sum of kernel\|CUDAFunctor_add\|upsample_fwd\|upsample_bwd\|conv_tc_fwd_2cta_kernel<2, 1>\|conv_tc_fwd_2cta_kernel<0, 1>\|spatial_broadcast" gpurun_out/r2u_prof_$m.log | cut -c1-150; done
B="python bench.py --no-gpu-baseline --no-fit --no-classifier --no-cpu-baseline --no-augment --steps 12 --warmup 3"
TAILN=1 stage r2u_bench_on 300 $B
CERVIX_GRAD_CHAIN=0 TAILN=1 stage r2u_bench_off 300 $B
TAILN=1 stage r2u_bench_on2 300 $B
python - <<'PY'
import json
for n in ("on","off","on2"):
    l=[x for x in open("gpurun_out/r2u_bench_%s.log"%n) if x.startswith("{")][-1]
    d=json.loads(l); print(n, d["value"], d["ms_per_step"])
PY
