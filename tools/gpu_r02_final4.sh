#!/bin/bash
# last validation of the round on the final build: whole GPU suite, smoke(), both bench arms as the driver runs them
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02f_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r02f_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02f_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02f_smoke.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02f_ref.json 2> gpurun_out/r02f_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02f_ours.json 2> gpurun_out/r02f_ours.err; echo "ours rc=$?"
python - <<PY
import json
for f in ("r02f_ref","r02f_ours"):
    o=json.loads([l for l in open("gpurun_out/%s.json"%f) if l.startswith("{")][-1]); print(f,{k:o.get(k) for k in ("value","ms_per_step","gpu_launches")}, o["e2e"]["value"], (o.get("one_batch_at_a_time") or {}).get("ms_per_step"))
PY
