#!/bin/bash
# Run on the GPU box via gpurun: kernel parity, tcgen05 probe, model parity, short bench.
# Every stage has its own timeout and log under gpurun_out/ so one failure does not hide the rest.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
python - <<'PY' > gpurun_out/env.txt 2>&1
import os, torch
print("torch", torch.__version__, "cuda", torch.cuda.is_available(), torch.cuda.get_device_name(0))
print("cpus", os.cpu_count())
PY
stage() { name=$1; shift; echo "=== $name" ; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log; }
stage kernels 600 python -m pytest tests/test_kernels_gpu.py -q -x --no-header -p no:cacheprovider
TAILN=30 stage tc_probe 900 python tools/tc_probe.py
CERVIX_DISABLE_TC=1 stage model_simt 900 python -m pytest tests/test_model_gpu.py -q --no-header -p no:cacheprovider -k "fp32 or bf16_eval"
stage tc_tests 600 python -m pytest tests/test_tc_gpu.py -q --no-header -p no:cacheprovider
stage model_tc 900 python -m pytest tests/test_model_gpu.py -q --no-header -p no:cacheprovider
stage smoke 600 python __graft_entry__.py --smoke
stage bench_small 900 python bench.py --steps 3 --warmup 2 --batch 8 --no-cpu-baseline
