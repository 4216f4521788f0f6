#!/bin/bash
set -u
mkdir -p gpurun_out
stage() { name=$1; shift; echo "=== $name" ; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log; }
TAILN=5 stage r2b_gpu_tests 1500 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider
TAILN=8 stage r2b_determinism 900 python tools/determinism_probe.py --cases f32:64:4,bf16:64:4,bf16:128:4,bf16:512:32
TAILN=2 stage r2b_ref_gpu_fp16 900 python oracle/ref_runner.py --device cuda --batch 32 --steps 6 --warmup 3 --fp16
TAILN=2 stage r2b_ref_gpu_fp16_cl 900 python oracle/ref_runner.py --device cuda --batch 32 --steps 6 --warmup 3 --fp16 --channels-last
TAILN=2 stage r2b_ref_gpu_fp32 900 python oracle/ref_runner.py --device cuda --batch 32 --steps 4 --warmup 2
TAILN=50 stage r2b_profile_step 900 python tools/profile_step.py --batch 32 --steps 2
