#!/bin/bash
# final 1-GPU records of the round: both arms as the driver runs them, a 200-step line, launch table, timeline, ncu graph-node traffic
mkdir -p gpurun_out
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02q_ref.json 2> gpurun_out/r02q_ref.err; echo "ref rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/r02q_ours.json 2> gpurun_out/r02q_ours.err; echo "ours rc=$?"
python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r02q_ours200.json 2> gpurun_out/r02q_ours200.err; echo "ours200 rc=$?"
python tools/launch_table.py 100 throughput > gpurun_out/r02q_launch_table.txt 2>&1
python tools/pipeline_timeline.py 8 16 > gpurun_out/r02q_pipeline_timeline.txt 2> gpurun_out/r02q_tl.err
python tools/pipeline_timeline.py 1 4 > gpurun_out/r02q_single_timeline.txt 2>> gpurun_out/r02q_tl.err
CMD="python bench.py --steps 1 --warmup 3 --pipeline 1 --no-cpu-baseline --no-latency-leg"
$CMD > gpurun_out/r02q_plain.log 2>&1 && ncu --graph-profiling node --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02q_graph_nodes.csv $CMD > gpurun_out/r02q_ncu.log 2>&1; echo "ncu rc=$?"
python - <<PY
import json
for f in ("r02q_ref","r02q_ours","r02q_ours200"):
    try:
        o=json.loads([l for l in open("gpurun_out/%s.json"%f) if l.startswith("{")][-1]); print(f,{k:o.get(k) for k in ("value","ms_per_step","e2e","e2e_fp32_image","one_batch_at_a_time","strict_fp32","cpu_baseline")}); 
        if "roofline" in o: print({k:o["roofline"].get(k) for k in ("achieved","frac","tensor_pipe_frac","traffic")})
    except Exception as e: print(f,"failed",e); print(open("gpurun_out/%s.err"%f).read()[-1500:])
PY
