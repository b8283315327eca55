#!/bin/bash
# prefix-FPS shortcut: parity, then the bench line (both legs) and the single-batch timeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fps_prefix_gpu.py tests/test_ops_gpu.py tests/test_backbone_gpu.py tests/test_reference_python_gpu.py -x -q -m gpu > gpurun_out/r02u_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02u_tests.log
timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r02u_bench.json 2> gpurun_out/r02u_bench.err; echo "bench rc=$?"
python - <<PY
import json
o=json.loads([l for l in open("gpurun_out/r02u_bench.json") if l.startswith("{")][-1])
print({k:o.get(k) for k in ("value","ms_per_step","e2e","one_batch_at_a_time","gpu_launches")})
print([ (r["op"],r["frac"]) for r in o.get("op_rooflines",[])])
PY
timeout 300 python tools/pipeline_timeline.py 1 > gpurun_out/r02u_single_timeline.txt 2> gpurun_out/r02u_single_timeline.err; head -8 gpurun_out/r02u_single_timeline.txt
