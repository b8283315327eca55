#!/bin/bash
# pipeline depth vs the driver's 20-step run; training step and rcnn_online with the new gradient / channel-major kernels
mkdir -p gpurun_out
for d in 8 12 16 24; do
  timeout 300 python bench.py --pipeline $d --steps 20 --warmup 5 --no-cpu-baseline --no-latency-leg > gpurun_out/r02v_p${d}_20.json 2> gpurun_out/r02v_p${d}_20.err; echo "depth $d steps 20 rc=$?"
done
for d in 12 16; do
  timeout 300 python bench.py --pipeline $d --steps 200 --warmup 5 --no-cpu-baseline --no-latency-leg > gpurun_out/r02v_p${d}_200.json 2> gpurun_out/r02v_p${d}_200.err; echo "depth $d steps 200 rc=$?"
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r02v_p*.json")):
    try:
        o=json.loads([l for l in open(f) if l.startswith("{")][-1]); print(f, o["value"], o["ms_per_step"], o["e2e"]["value"], o.get("e2e_fp32_image",{}).get("value"))
    except Exception as e: print(f,"failed",e)
PY
timeout 600 python bench.py --mode train --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02v_train.json 2> gpurun_out/r02v_train.err; echo "train rc=$?"; tail -c 600 gpurun_out/r02v_train.json
timeout 600 python tools/rcnn_online.py --steps 20 --warmup 3 > gpurun_out/r02v_rcnn_online.txt 2> gpurun_out/r02v_rcnn_online.err; echo "rcnn rc=$?"; cat gpurun_out/r02v_rcnn_online.txt | cut -c1-400
