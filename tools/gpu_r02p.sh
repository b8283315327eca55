#!/bin/bash
mkdir -p gpurun_out
for p in 0 1 0 1; do
EPNET_FPS_PRIORITY=$p timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r02p_prio${p}.json 2> gpurun_out/r02p_prio${p}.err; echo "prio $p rc=$?"
python - <<PY
import json
o=json.loads([l for l in open("gpurun_out/r02p_prio${p}.json") if l.startswith("{")][-1]); print("prio ${p}", o["value"], o["ms_per_step"], o["e2e"]["value"], o["one_batch_at_a_time"]["ms_per_step"])
PY
done
