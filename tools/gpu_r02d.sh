#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02d_tests.log 2>&1; echo "tests rc=$?"
grep -E "passed|failed|FAILED|Error" gpurun_out/r02d_tests.log | tail -15
timeout 300 python tools/sanitize_zoo.py > gpurun_out/r02d_zoo_plain.log 2>&1; echo "zoo plain rc=$?"; tail -3 gpurun_out/r02d_zoo_plain.log
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_zoo.py > gpurun_out/r02d_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -8 gpurun_out/r02d_memcheck.log
