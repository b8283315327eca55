#!/bin/bash
# first convolution as the restructured SIMT kernel: parity, launch table, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_tma_gpu.py tests/test_backbone_gpu.py tests/test_reference_python_gpu.py tests/test_f16_guard_gpu.py tests/test_image_prep_gpu.py -x -q -m gpu > gpurun_out/r02k_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02k_tests.log
timeout 300 python tools/launch_table.py > gpurun_out/r02k_launch_table.txt 2>&1; grep -n "conv3x3\|^total" gpurun_out/r02k_launch_table.txt | head -14
timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err; echo "bench rc=$?"
python - <<PY
import json
o=json.loads([l for l in open("gpurun_out/r02k_bench.json") if l.startswith("{")][-1])
print({k:o.get(k) for k in ("value","ms_per_step","gpu_launches")}, o["e2e"]["value"], o["one_batch_at_a_time"]["value"], o["one_batch_at_a_time"]["ms_per_step"], o["roofline"]["frac"])
PY
