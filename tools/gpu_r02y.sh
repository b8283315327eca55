#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_gemm_tma_gpu.py tests/test_backbone_gpu.py tests/test_reference_python_gpu.py -x -q -m gpu > gpurun_out/r02y_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02y_tests.log
timeout 300 python tools/launch_table.py > gpurun_out/r02y_launch_table.txt 2>&1; head -12 gpurun_out/r02y_launch_table.txt; grep -E "\((128|512|2048), (2048|1536|1024|768|512)," gpurun_out/r02y_launch_table.txt
timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r02y_bench.json 2> gpurun_out/r02y_bench.err; echo "bench rc=$?"
python - <<PY
import json
o=json.loads([l for l in open("gpurun_out/r02y_bench.json") if l.startswith("{")][-1])
print({k:o.get(k) for k in ("value","ms_per_step","gpu_launches")}, o["e2e"]["value"], o["one_batch_at_a_time"]["value"], o["one_batch_at_a_time"]["ms_per_step"], o["roofline"]["frac"])
PY
