#!/bin/bash
# ncu launch list (gpu__time_duration per launch) of one eager bench step; runs only after the same command exited 0
set -u
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-classifier --no-graph > gpurun_out/bench_nograph_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 4200 -c 1500 --csv --log-file gpurun_out/launches_b32.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-classifier --no-graph > gpurun_out/ncu_launches.log 2>&1
echo "rc=$? (ncu launches)"; tail -1 gpurun_out/ncu_launches.log | cut -c1-200
