#!/bin/bash
set -u
mkdir -p gpurun_out
stage() { name=$1; shift; echo "=== $name" ; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n "${TAILN:-6}" gpurun_out/$name.log; }
TAILN=60 stage prof_b32 600 python tools/profile_step.py --batch 32 --steps 2
stage bench_b32 900 python bench.py --steps 5 --warmup 3
stage bench_ref 400 python bench.py --impl reference --steps 3 --warmup 1
stage bench_plain_b8 600 python bench.py --steps 1 --warmup 1 --batch 8 --no-cpu-baseline
if [ $? -eq 0 ]; then
TAILN=3 stage ncu_launches 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_b8.csv python bench.py --steps 1 --warmup 1 --batch 8 --no-cpu-baseline
fi
