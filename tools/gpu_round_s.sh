#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_patch_encoder.py -q -x --no-header -p no:cacheprovider 2>&1 | tail -2
python tools/profile_encoder.py 2>&1 | grep -A5 "256 patches" | cut -c1-150
