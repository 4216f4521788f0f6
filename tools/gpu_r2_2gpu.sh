#!/bin/bash
set -u
mkdir -p gpurun_out
stage() { name=$1; shift; echo "=== $name" ; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
TAILN=1 stage r2m_bench2 500 $TR bench.py --gpus 2 --steps 12 --warmup 3 --no-classifier
TAILN=6 stage r2m_ddp_tests 300 python -m pytest tests/test_ddp_gpu.py -q --no-header -p no:cacheprovider
