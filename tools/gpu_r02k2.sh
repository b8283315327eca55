#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_tma_gpu.py -x -q -m gpu > gpurun_out/r02k_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02k_tests.log
timeout 120 python tools/first_conv_probe.py 2>&1 | tee gpurun_out/r02k_probe.txt
