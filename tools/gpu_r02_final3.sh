#!/bin/bash
# final 2-GPU records (weak scaling, as the driver launches them) + the 2-GPU column of the config-5 sweep with the final kernels
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521"
timeout 400 $TR bench.py --gpus 2 --steps 20 --warmup 5 --impl reference > gpurun_out/r02d_ref2.json 2> gpurun_out/r02d_ref2.err; echo "ref2 rc=$?"
timeout 400 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02d_ours2.json 2> gpurun_out/r02d_ours2.err; echo "ours2 rc=$?"
timeout 600 $TR tests/perf/op_sweep.py gpurun_out/r02d_op_sweep_2gpu.json > gpurun_out/r02d_sweep2.log 2>&1; echo "sweep2 rc=$?"
timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_backward_gpu.py -x -q -m gpu > gpurun_out/r02d_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r02d_tests.log
timeout 600 python tests/perf/op_sweep.py gpurun_out/r02d_op_sweep.json > gpurun_out/r02d_op_sweep.log 2>&1; echo "sweep1 rc=$?"
python - <<PY
import json
for f in ("r02d_ref2","r02d_ours2"):
    try:
        o=json.loads([l for l in open("gpurun_out/%s.json"%f) if l.startswith("{")][-1]); print(f,{k:o.get(k) for k in ("value","n_gpus","ms_per_step","e2e","e2e_fp32_image","strict_fp32")})
    except Exception as e: print(f,"failed",e); print(open("gpurun_out/%s.err"%f).read()[-1200:])
PY
du -sh gpurun_out
