#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_fused_gpu.py -q -x --no-header -p no:cacheprovider 2>&1 | tail -2
python tools/bench_conv_shapes.py 2>&1 | tail -28
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-classifier 2>&1 | tail -1 | cut -c1-200
