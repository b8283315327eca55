#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/rcnn_online.py --steps 20 --warmup 3 > gpurun_out/r02g_rcnn_online.txt 2> gpurun_out/r02g_rcnn_online.err; echo "rcnn rc=$?"
cat gpurun_out/r02g_rcnn_online.txt; tail -5 gpurun_out/r02g_rcnn_online.err
timeout 600 python bench.py --mode train --impl reference --steps 10 --warmup 3 > gpurun_out/r02g_train_ref.json 2> gpurun_out/r02g_train_ref.err; echo "train ref rc=$?"
timeout 600 python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/r02g_train_ours.json 2> gpurun_out/r02g_train_ours.err; echo "train ours rc=$?"
python - <<PY
import json
for f in ("r02g_train_ref","r02g_train_ours"):
    try:
        o=json.load(open("gpurun_out/%s.json"%f)); print(f,{k:o.get(k) for k in ("value","ms_per_step","e2e","strict_fp32")})
    except Exception as e: print(f,"failed",e); print(open("gpurun_out/%s.err"%f).read()[-1500:])
PY
