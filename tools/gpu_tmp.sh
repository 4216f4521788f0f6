#!/bin/bash
set -u
timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_fused_gpu.py tests/test_patch_encoder.py tests/test_kernels_gpu.py -q -x --no-header -p no:cacheprovider 2>&1 | tail -2
CERVIX_TC_1CTA=1 timeout 600 python -m pytest tests/test_tc_gpu.py -q -x --no-header -p no:cacheprovider 2>&1 | tail -1
python tools/profile_encoder.py 2>&1 | grep -A5 "256 patches" | cut -c1-140
