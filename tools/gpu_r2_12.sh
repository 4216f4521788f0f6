#!/bin/bash
# 2-GPU call: engine + DDP tests, then the data-parallel bench with the two-graph split backward and with the one-graph step
set -u
mkdir -p gpurun_out
stage() { name=$1; shift; echo "=== $name" ; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log; }
TAILN=8 stage r2o_engine_tests 900 python -m pytest tests/test_engine_gpu.py tests/test_ddp_gpu.py -m gpu -q --no-header -p no:cacheprovider -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
TAILN=2 stage r2o_bench2_split 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-gpu-baseline --no-fit --no-strong
CERVIX_SPLIT_BACKWARD=0 TAILN=2 stage r2o_bench2_onegraph 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-gpu-baseline --no-fit --no-strong
