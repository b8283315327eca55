#!/bin/bash
mkdir -p gpurun_out
for r in 1 2 3; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-latency-leg > gpurun_out/r02q2_${r}.json 2> gpurun_out/r02q2_${r}.err
python - <<PY
import json
o=json.loads([l for l in open("gpurun_out/r02q2_${r}.json") if l.startswith("{")][-1]); print("run ${r} steps 20:", o["value"], o["e2e"]["value"], o["e2e_fp32_image"]["value"])
PY
done
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-latency-leg > gpurun_out/r02q2_200.json 2> gpurun_out/r02q2_200.err
python - <<PY
import json
o=json.loads([l for l in open("gpurun_out/r02q2_200.json") if l.startswith("{")][-1]); print("steps 200:", o["value"], o["e2e"]["value"], o["e2e_fp32_image"]["value"])
PY
