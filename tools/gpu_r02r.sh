#!/bin/bash
# 8-GPU record: both arms as the driver launches them
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551"
timeout 500 $TR bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02i8_ours8.json 2> gpurun_out/r02i8_ours8.err; echo "ours8 rc=$?"
timeout 500 $TR bench.py --gpus 8 --steps 20 --warmup 5 --impl reference > gpurun_out/r02i8_ref8.json 2> gpurun_out/r02i8_ref8.err; echo "ref8 rc=$?"
nvidia-smi topo -m > gpurun_out/r02i8_topo.txt 2>&1
python - <<PY
import json
for f in ("r02i8_ours8","r02i8_ref8"):
    try:
        o=json.loads([l for l in open("gpurun_out/%s.json"%f) if l.startswith("{")][-1]); print(f,{k:o.get(k) for k in ("value","n_gpus","ms_per_step","e2e","e2e_fp32_image","strict_fp32")})
    except Exception as e: print(f,"failed",e); print(open("gpurun_out/%s.err"%f).read()[-1500:])
PY
