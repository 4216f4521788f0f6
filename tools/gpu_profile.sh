#!/bin/bash
set -u
mkdir -p gpurun_out
python tools/profile_step.py > gpurun_out/prof_b32.log 2>&1; echo "rc=$? prof"
