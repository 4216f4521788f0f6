#!/bin/bash
# ncu evidence for the judged numbers: full-metric captures of the roofline kernel (decoder 3x3 conv forward), the
# CTA-pair weight gradient and the middle-flow pointwise conv.  Each ncu step runs only after the same command exited 0
# without ncu.  (The launch list of the bench command is tools/gpu_ncu_launches.sh: ~8 min of box time.)
set -u
mkdir -p gpurun_out
cap() {  # name kernel-regex args...
  name=$1; regex=$2; shift 2
  python tools/run_conv_once.py "$@" > gpurun_out/plain_$name.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$regex -s 2 -c 1 -o gpurun_out/ncu_$name \
      python tools/run_conv_once.py "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "rc=$? ($name)"
}
cap conv_fwd_r01v6 conv_tc_fwd 32 fwd
cap conv_wgrad_r01v6 conv_tc_wgrad 32 wgrad
cap pw728_fwd_r01v6 conv_tc_fwd 32 fwd 728 728 1 32
cap pw728_wgrad_r01v6 conv_tc_wgrad 32 wgrad 728 728 1 32
