#!/bin/bash
# final records of round 2 on one B200: the whole GPU suite, smoke(), both bench arms as the driver runs them, a 200-step line, the
# config-5 sweep with the final kernels, launch table, timelines, ncu graph-node traffic of one replay
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02h_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02h_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02h_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02h_smoke.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02h_ref.json 2> gpurun_out/r02h_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02h_ours.json 2> gpurun_out/r02h_ours.err; echo "ours rc=$?"
timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r02h_ours200.json 2> gpurun_out/r02h_ours200.err; echo "ours200 rc=$?"
timeout 300 python tools/launch_table.py 100 throughput > gpurun_out/r02h_launch_table.txt 2>&1
timeout 300 python tools/pipeline_timeline.py 8 16 > gpurun_out/r02h_pipeline_timeline.txt 2> gpurun_out/r02h_tl.err
timeout 300 python tools/pipeline_timeline.py 1 4 > gpurun_out/r02h_single_timeline.txt 2>> gpurun_out/r02h_tl.err
CMD="python bench.py --steps 1 --warmup 3 --pipeline 1 --no-cpu-baseline --no-latency-leg"
timeout 300 $CMD > gpurun_out/r02h_plain.log 2>&1 && timeout 600 ncu --graph-profiling node --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02h_graph_nodes.csv $CMD > gpurun_out/r02h_ncu.log 2>&1; echo "ncu rc=$?"
python - <<PY
import json
for f in ("r02h_ref","r02h_ours","r02h_ours200"):
    try:
        o=json.loads([l for l in open("gpurun_out/%s.json"%f) if l.startswith("{")][-1]); print(f,{k:o.get(k) for k in ("value","ms_per_step","e2e","e2e_fp32_image","one_batch_at_a_time","strict_fp32","cpu_baseline")});
        if "roofline" in o: print({k:o["roofline"].get(k) for k in ("achieved","frac","tensor_pipe_frac","traffic")})
    except Exception as e: print(f,"failed",e); print(open("gpurun_out/%s.err"%f).read()[-1500:])
PY
du -sh gpurun_out
