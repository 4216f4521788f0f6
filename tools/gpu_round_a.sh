#!/bin/bash
# conv shape table + GPU suite + ncu of the middle-flow pointwise conv (728->728 @32x32, batch 32)
set -u
mkdir -p gpurun_out
stage() { name=$1; shift; echo "=== $name" ; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log; }
TAILN=40 stage conv_shapes 900 python tools/bench_conv_shapes.py
TAILN=40 stage conv_shapes_noflush 900 python tools/bench_conv_shapes.py --no-flush
stage gpu_tests_all 1500 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider
python tools/run_conv_once.py 32 fwd 728 728 1 32 > gpurun_out/pw_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_tc_fwd -s 2 -c 1 -o gpurun_out/ncu_pw728_fwd \
    python tools/run_conv_once.py 32 fwd 728 728 1 32 > gpurun_out/ncu_pw728_fwd.log 2>&1
echo "rc=$? ncu pw fwd"
ncu --set full --clock-control none --import-source on -k regex:conv_tc_wgrad -s 2 -c 1 -o gpurun_out/ncu_pw728_wgrad \
    python tools/run_conv_once.py 32 wgrad 728 728 1 32 > gpurun_out/ncu_pw728_wgrad.log 2>&1
echo "rc=$? ncu pw wgrad"
