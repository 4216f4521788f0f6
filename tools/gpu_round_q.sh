#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fused_gpu.py tests/test_tc_gpu.py -q -x --no-header -p no:cacheprovider 2>&1 | tail -3
python tools/bench_fused.py --graph --only conv_ 2>&1 | tail -5
CERVIX_BN1_IN_DGRAD=1 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-classifier 2>&1 | tail -1 | cut -c1-200
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-classifier 2>&1 | tail -1 | cut -c1-200
