#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fused_gpu.py tests/test_patch_encoder.py -q -x --no-header -p no:cacheprovider 2>&1 | tail -2
python tools/profile_encoder.py 2>&1 | grep -A4 "256 patches" | cut -c1-150
python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['value'],1), {k:(round(v,2) if isinstance(v,float) else v) for k,v in d['classifier'].items() if k!='workload'})"
