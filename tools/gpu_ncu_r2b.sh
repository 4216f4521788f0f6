#!/bin/bash
# round-2 evidence, final build: event-timed launches of the kernel families touched late in the round, one `ncu --set full`
# capture of each new kernel (summarised on the box), and the ncu launch list of one eager bench step.
set -u
mkdir -p gpurun_out
timeout 600 python tools/run_families_once.py > gpurun_out/r02_families_v3.log 2>&1; echo "rc=$? (plain)"; tail -32 gpurun_out/r02_families_v3.log
: > gpurun_out/r02_ncu_summary_v3.txt
cap() { # name, kernel regex, --only filter
  timeout 200 ncu --set full --clock-control none -k regex:"$2" -s ${4:-2} -c 1 -o /tmp/ncu_r02_$1 \
      python tools/run_families_once.py --only "$3" --reps 1 > /tmp/ncu_r02_$1.log 2>&1
  echo "rc=$? (ncu $1)"
  python tools/ncu_summary.py /tmp/ncu_r02_$1.ncu-rep >> gpurun_out/r02_ncu_summary_v3.txt 2>&1
  rm -f /tmp/ncu_r02_$1.ncu-rep
}
cap upsample_concat "upsample_fwd_kernel" "upsample_concat"
cap dropout "dropout_kernel" "dropout_fwd"
cap aug_rows "aug_resize_rows_kernel" "augment batch"
cap aug_compose "aug_compose_kernel" "augment batch"
cap aug_rotate "aug_rotate_jitter_kernel" "augment batch"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file /tmp/r02_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-graph --no-classifier --no-cpu-baseline --no-gpu-baseline --no-fit --no-augment > gpurun_out/r02_launches_v3.log 2>&1
echo "rc=$? (launch list)"
gzip -c /tmp/r02_launches.csv > gpurun_out/r02_launches_v3.csv.gz
ls -la gpurun_out | tail -8
