#!/bin/bash
# quick GPU check after a kernel change: whole-backbone parity tests, the bench line, the pipelined timeline
mkdir -p gpurun_out
T=${1:-quick}
timeout 600 python -m pytest tests/test_gemm_tma_gpu.py tests/test_backbone_gpu.py tests/test_reference_python_gpu.py tests/test_golden_gpu.py tests/test_f16_guard_gpu.py -m gpu -q -s > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"
grep -E "passed|failed|FAILED|worst|vs module" gpurun_out/${T}_tests.log | tail -20
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python - <<PY
import json
try:
    o=json.load(open("gpurun_out/${T}_bench.json"))
    print({k:o.get(k) for k in ("value","ms_per_step","e2e","e2e_fp32_image","one_batch_at_a_time")})
    print({k:o["roofline"].get(k) for k in ("achieved","frac","tensor_pipe_frac","isolated")})
    for k in o["kernel_breakdown"][:12]: print(k)
except Exception as e: print("no bench line", e); print(open("gpurun_out/${T}_bench.err").read()[-2000:])
PY
timeout 300 python tools/pipeline_timeline.py 8 16 > gpurun_out/${T}_timeline.txt 2> gpurun_out/${T}_timeline.err; head -22 gpurun_out/${T}_timeline.txt
