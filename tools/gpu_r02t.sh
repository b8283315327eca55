#!/bin/bash
# channel-major ops through the point-major scratch copy: parity, then timing against the staged path
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_fused_gpu.py tests/test_backward_gpu.py -x -q -m gpu > gpurun_out/r02t_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02t_tests.log
EPNET_CM_ROWS=transposed timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "interpolate or group" > gpurun_out/r02t_tests_forced.log 2>&1; echo "forced tests rc=$?"; tail -3 gpurun_out/r02t_tests_forced.log
timeout 300 python tools/op_roofline_probe.py auto > gpurun_out/r02t_ops_auto.txt 2>&1; cat gpurun_out/r02t_ops_auto.txt
EPNET_CM_ROWS=staged timeout 300 python tools/op_roofline_probe.py staged > gpurun_out/r02t_ops_staged.txt 2>&1; grep -E "group_points|gather_points|three_interpolate " gpurun_out/r02t_ops_staged.txt
EPNET_CM_ROWS=transposed timeout 300 python tools/op_roofline_probe.py transposed > gpurun_out/r02t_ops_transposed.txt 2>&1; grep -E "group_points|gather_points|three_interpolate " gpurun_out/r02t_ops_transposed.txt
