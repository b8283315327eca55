#!/bin/bash
# 2-GPU call: bench both arms under torchrun (as the driver launches them), training step 2 GPUs (DDP all-reduce), op sweep 2 GPUs
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 --impl reference > gpurun_out/r02j_ref2.json 2> gpurun_out/r02j_ref2.err; echo "ref2 rc=$?"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02j_ours2.json 2> gpurun_out/r02j_ours2.err; echo "ours2 rc=$?"
timeout 600 $TR bench.py --gpus 2 --mode train --steps 10 --warmup 3 --impl reference > gpurun_out/r02j_train_ref2.json 2> gpurun_out/r02j_train_ref2.err; echo "train ref2 rc=$?"
timeout 600 $TR bench.py --gpus 2 --mode train --steps 10 --warmup 3 > gpurun_out/r02j_train_ours2.json 2> gpurun_out/r02j_train_ours2.err; echo "train ours2 rc=$?"
timeout 600 $TR bench.py --gpus 2 --mode train --steps 10 --warmup 3 --tf32 1 > gpurun_out/r02j_train_ours2_tf32.json 2> gpurun_out/r02j_train_ours2_tf32.err; echo "train ours2 tf32 rc=$?"
timeout 900 $TR tests/perf/op_sweep.py gpurun_out/r02_op_sweep_2gpu.json > gpurun_out/r02j_sweep2.log 2>&1; echo "sweep2 rc=$?"
python - <<PY
import json
for f in ("r02j_ref2","r02j_ours2","r02j_train_ref2","r02j_train_ours2","r02j_train_ours2_tf32"):
    try:
        o=json.loads([l for l in open("gpurun_out/%s.json"%f) if l.startswith("{")][-1]); print(f,{k:o.get(k) for k in ("value","n_gpus","ms_per_step","e2e","e2e_uint8_image","strict_fp32")})
    except Exception as e: print(f,"failed",e); print(open("gpurun_out/%s.err"%f).read()[-1200:])
PY
