#!/bin/bash
set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 300 python -m pytest "$@" -q --no-header -p no:cacheprovider 2>&1 | tail -3 | cut -c1-300; }
run whole_file tests/test_fusion_gpu.py -x
run only_graph tests/test_fusion_gpu.py -k "graph_captured"
run dropout_then_graph tests/test_fusion_gpu.py -k "dropout or graph_captured"
run golden_then_graph tests/test_fusion_gpu.py -k "golden or graph_captured"
run batch_then_graph tests/test_fusion_gpu.py -k "batch_equals or graph_captured"
run age_then_graph tests/test_fusion_gpu.py -k "age_node or graph_captured"
run engine_then_graph tests/test_engine_gpu.py tests/test_fusion_gpu.py -k "test_engine_gpu or graph_captured"
run fused_then_graph tests/test_fused_gpu.py tests/test_fusion_gpu.py -k "test_fused_gpu or graph_captured"
run baseline_then_graph tests/test_baseline_configs_gpu.py tests/test_fusion_gpu.py -k "config1 or config0 or graph_captured"
