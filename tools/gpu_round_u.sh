#!/bin/bash
set -u
mkdir -p gpurun_out
CERVIX_TC_PAIRS=2 python tools/bench_conv_shapes.py > gpurun_out/conv_shapes_p2.log 2>&1
CERVIX_TC_PAIRS=4 python tools/bench_conv_shapes.py > gpurun_out/conv_shapes_p4.log 2>&1
python tools/bench_conv_shapes.py > gpurun_out/conv_shapes_p1.log 2>&1
echo done
