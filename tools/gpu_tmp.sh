#!/bin/bash
set -u
timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_fused_gpu.py tests/test_patch_encoder.py -q -x --no-header -p no:cacheprovider 2>&1 | tail -2
python tools/profile_encoder.py 2>&1 | grep -A3 "256 patches" | cut -c1-140
python tools/bench_fused.py --graph --only conv_ 2>&1 | tail -5
python tools/bench_conv_shapes.py 2>&1 | grep -E "conv2|b1.pw|mid.pw|cat_conv|total"
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-classifier 2>&1 | tail -1 | cut -c1-200
