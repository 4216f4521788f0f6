#!/bin/bash
# one quick confirmation pass: the tests named in $TESTS (default: the kernel files), then a bench line
set -u
mkdir -p gpurun_out
stage() { name=$1; shift; echo "=== $name" ; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log; }
stage step_tests 900 python -m pytest ${TESTS:-tests/test_kernels_gpu.py tests/test_model_gpu.py} -q --no-header -p no:cacheprovider -m gpu -x
TAILN=1 stage step_bench 900 python bench.py ${BENCH_ARGS:---no-cpu-baseline --no-classifier}
python - <<'PY'
import json
l=[x for x in open('gpurun_out/step_bench.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('VALUE %.1f img/s  %.2f ms/step  e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']))
PY
