#!/bin/bash
mkdir -p gpurun_out
for d in 8 6 5 7 8 6; do
  timeout 300 python bench.py --pipeline $d --steps 20 --warmup 5 --no-cpu-baseline --no-latency-leg > gpurun_out/r02n_p${d}.json 2> gpurun_out/r02n_p${d}.err
  python - <<PY
import json
o=json.loads([l for l in open("gpurun_out/r02n_p${d}.json") if l.startswith("{")][-1]); print("depth ${d} steps 20:", o["value"], o["e2e"]["value"], o["e2e_fp32_image"]["value"])
PY
done
for d in 6; do
  timeout 300 python bench.py --pipeline $d --steps 200 --warmup 5 --no-cpu-baseline --no-latency-leg > gpurun_out/r02n_p${d}_200.json 2> gpurun_out/r02n_p${d}_200.err
  python - <<PY
import json
o=json.loads([l for l in open("gpurun_out/r02n_p${d}_200.json") if l.startswith("{")][-1]); print("depth ${d} steps 200:", o["value"], o["e2e"]["value"], o["e2e_fp32_image"]["value"])
PY
done
