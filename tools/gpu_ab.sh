#!/bin/bash
# A/B of an environment switch: bash tools/gpu_ab.sh VAR  [A B] -> bench with VAR=A (default 0) and VAR=B (default 1), twice each, interleaved
set -u
mkdir -p gpurun_out
for rep in 1 2; do for v in ${2:-0} ${3:-1}; do
  env $1=$v python bench.py --no-cpu-baseline --no-classifier --steps 12 > gpurun_out/ab_${v}_${rep}.log 2>&1
  python - "$1" "$v" gpurun_out/ab_${v}_${rep}.log <<'PY'
import json,sys
l=[x for x in open(sys.argv[3]) if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('%s=%s: %.1f img/s  %.2f ms/step' % (sys.argv[1], sys.argv[2], d['value'], d['ms_per_step']))
else:
    print(sys.argv[1], sys.argv[2], 'FAILED'); print(open(sys.argv[3]).read()[-800:])
PY
done; done
