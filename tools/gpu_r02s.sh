#!/bin/bash
# staged-rows rewrite (prefetch, 1024-thread CTAs, parts mode), vectorised NCHW taps, vector-RED probe
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_fused_gpu.py -x -q -m gpu > gpurun_out/r02s_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02s_tests.log
timeout 60 tools/probes/_bin/red_probe > gpurun_out/r02s_red_probe.txt 2>&1; cat gpurun_out/r02s_red_probe.txt
timeout 300 python tools/op_roofline_probe.py parts > gpurun_out/r02s_ops_parts.txt 2>&1; cat gpurun_out/r02s_ops_parts.txt
EPNET_STAGED_LONG_ROWS=partial timeout 300 python tools/op_roofline_probe.py partial > gpurun_out/r02s_ops_partial.txt 2>&1; grep -E "group_points|gather_points" gpurun_out/r02s_ops_partial.txt
