#!/bin/bash
# full GPU suite (no -x: every failure is listed), log under gpurun_out/
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/${1:-tests}.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/${1:-tests}.log
grep -E "passed|failed|FAILED|Error" gpurun_out/${1:-tests}.log | tail -30
