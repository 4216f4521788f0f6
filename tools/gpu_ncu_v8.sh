#!/bin/bash
# v8 evidence: event-timed micro-benchmarks of the new kernels, one `ncu --set full` capture of each (after the same
# command exited 0 without ncu), the step breakdown and the ncu launch list of one eager bench step.
set -u
mkdir -p gpurun_out
python tools/run_new_kernels_once.py > gpurun_out/new_kernels_v8.log 2>&1; echo "rc=$? (plain)"; cat gpurun_out/new_kernels_v8.log
for k in dw_s2_dgrad stem_wgrad_mma upsample_to_nchw_bwd split_patches bn_bwd_apply "colreduce.*BnBwdF"; do
  n=$(echo $k | tr -cd 'a-z0-9_')
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/ncu_v8_$n \
      python tools/run_new_kernels_once.py > gpurun_out/ncu_v8_$n.log 2>&1
  echo "rc=$? (ncu $k)"
done
python tools/ncu_summary.py gpurun_out/ncu_v8_*.ncu-rep > gpurun_out/ncu_v8_summary.txt 2>&1
python tools/profile_step.py > gpurun_out/prof_b32.log 2>&1; echo "rc=$? (profile_step)"
bash tools/gpu_ncu_launches.sh
