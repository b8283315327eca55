#!/bin/bash
# the driver's pair on one box with the final bench.py (host-side result ring twice the pipeline depth)
mkdir -p gpurun_out
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02j_ref.json 2> gpurun_out/r02j_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02j_ours.json 2> gpurun_out/r02j_ours.err; echo "ours rc=$?"
timeout 300 python -m pytest tests/test_reference_python_gpu.py tests/test_dist_cpu.py -x -q > gpurun_out/r02j_tests.log 2>&1; echo "tests rc=$?"; tail -1 gpurun_out/r02j_tests.log
python - <<PY
import json
for f in ("r02j_ref","r02j_ours"):
    o=json.loads([l for l in open("gpurun_out/%s.json"%f) if l.startswith("{")][-1]); print(f,{k:o.get(k) for k in ("value","ms_per_step","gpu_launches")}, o["e2e"]["value"], (o.get("e2e_fp32_image") or {}).get("value"), (o.get("one_batch_at_a_time") or {}).get("ms_per_step"))
PY
