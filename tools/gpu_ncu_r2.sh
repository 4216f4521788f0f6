#!/bin/bash
# round-2 evidence: event-timed launches of every kernel family, then ONE `ncu --set full` capture per family (after the
# plain run exited 0), each summarised on the box by tools/ncu_summary.py (the .ncu-rep files are too big to bring back),
# then the ncu launch list of one eager bench step.
set -u
mkdir -p gpurun_out
timeout 600 python tools/run_families_once.py > gpurun_out/r02_families.log 2>&1; echo "rc=$? (plain)"; cat gpurun_out/r02_families.log
: > gpurun_out/r02_ncu_summary.txt
cap() { # name, kernel regex, --only filter
  timeout 240 ncu --set full --clock-control none -k regex:"$2" -s 2 -c 1 -o /tmp/ncu_r02_$1 \
      python tools/run_families_once.py --only "$3" --reps 1 > /tmp/ncu_r02_$1.log 2>&1
  echo "rc=$? (ncu $1)"
  python tools/ncu_summary.py /tmp/ncu_r02_$1.ncu-rep >> gpurun_out/r02_ncu_summary.txt 2>&1
  rm -f /tmp/ncu_r02_$1.ncu-rep
}
cap dwf_bwd "dwf_bwd_kernel" "dwf_bwd affine"
cap dwf_bwd_addend "dwf_bwd_kernel" "dwf_bwd plain"
cap dwf_fwd "dwf_fwd_kernel" "dwf_fwd"
cap midpw_fwd "conv_tc_fwd_2cta" "conv_fwd_ex stats"
cap midpw_wgrad "conv_tc_wgrad_2cta" "conv_wgrad tc 728"
cap bnraw "colreduce_kernel" "bn_bwd_sums"
cap bn_bwd_affine "bn_bwd_affine_kernel" "bn_bwd_affine"
cap stats "colreduce_kernel" "bn_stats"
cap bn_apply "bn_apply_kernel" "bn_forward"
cap bn_bwd_apply "bn_bwd_apply_kernel" "bn_backward"
cap dw_dil2 "dw_" "dw_fwd dilation 2"
cap loss_stats "seg_loss_stats_kernel" "seg_loss_stats"
cap loss_grad "seg_loss_grad_kernel" "seg_loss_grad"
cap upsample_nchw_fwd "upsample_to_nchw_fwd" "upsample_to_nchw_fwd"
cap adam "adam_dev_kernel" "adam_dev"
cap gemm_grouped "gemm_grouped_kernel" "gemm_grouped"
cap enc_3x3 "conv_tc_fwd" "encoder conv_fwd_act 3x3"
# launch list of one eager bench step (cold-cache, serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file /tmp/r02_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-graph --no-classifier --no-cpu-baseline --no-gpu-baseline > gpurun_out/r02_launches.log 2>&1
echo "rc=$? (launch list)"
gzip -c /tmp/r02_launches.csv > gpurun_out/r02_launches.csv.gz
ls -la gpurun_out
