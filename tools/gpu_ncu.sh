#!/bin/bash
# ncu evidence for the judged numbers: (1) full-metric captures of the roofline kernel (decoder 3x3 conv forward), the
# CTA-pair weight gradient and the middle-flow pointwise conv, (2) launch list of the bench command.  Each ncu step
# runs only after the same command exited 0 without ncu.
set -u
mkdir -p gpurun_out
cap() {  # name kernel-regex args...
  name=$1; regex=$2; shift 2
  python tools/run_conv_once.py "$@" > gpurun_out/plain_$name.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$regex -s 2 -c 1 -o gpurun_out/ncu_$name \
      python tools/run_conv_once.py "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "rc=$? ($name)"
}
cap conv_fwd_r01v4 conv_tc_fwd 32 fwd
cap conv_wgrad_r01v4 conv_tc_wgrad 32 wgrad
cap pw728_fwd_r01v4 conv_tc_fwd 32 fwd 728 728 1 32
cap pw728_wgrad_r01v4 conv_tc_wgrad 32 wgrad 728 728 1 32
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-classifier --no-graph > gpurun_out/bench_nograph_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 5800 -c 2000 --csv --log-file gpurun_out/launches_b32.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-classifier --no-graph > gpurun_out/ncu_launches.log 2>&1
echo "rc=$? (ncu launches)"; tail -2 gpurun_out/ncu_launches.log | cut -c1-300
python tools/profile_step.py > gpurun_out/prof_b32.log 2>&1; echo "rc=$? prof"
