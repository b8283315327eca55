#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_sparse_tail_gpu.py tests/test_fused_gpu.py -x -q -m gpu > gpurun_out/r02o_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r02o_tests.log
timeout 120 python tools/small_m_probe.py 2>&1 | head -4
timeout 300 python tools/launch_table.py > gpurun_out/r02o_launch_table.txt 2>&1; head -8 gpurun_out/r02o_launch_table.txt
timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-latency-leg > gpurun_out/r02o_bench.json 2> gpurun_out/r02o_bench.err; echo "bench rc=$?"
python - <<PY
import json
o=json.loads([l for l in open("gpurun_out/r02o_bench.json") if l.startswith("{")][-1])
print({k:o.get(k) for k in ("value","ms_per_step")}, o["e2e"]["value"])
PY
