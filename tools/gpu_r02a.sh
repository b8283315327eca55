#!/bin/bash
# GPU call r02a: full GPU suite, both bench arms as the driver runs them, the pipelined timeline, ncu node list of one replay
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/r02a_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r02a_tests.log
tail -5 gpurun_out/r02a_tests.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02a_ref.json 2> gpurun_out/r02a_ref.err; echo "ref rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_ours.json 2> gpurun_out/r02a_ours.err; echo "ours rc=$?"
python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r02a_ours200.json 2> gpurun_out/r02a_ours200.err; echo "ours200 rc=$?"
python tools/pipeline_timeline.py 8 16 > gpurun_out/r02a_pipeline_timeline.txt 2> gpurun_out/r02a_pipeline_timeline.err; echo "timeline rc=$?"
python tools/pipeline_timeline.py 1 4 > gpurun_out/r02a_single_timeline.txt 2>> gpurun_out/r02a_pipeline_timeline.err
CMD="python bench.py --steps 1 --warmup 3 --pipeline 1 --no-cpu-baseline --no-latency-leg"
$CMD > gpurun_out/r02a_plain.log 2>&1 && ncu --graph-profiling node --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02a_graph_nodes.csv $CMD > gpurun_out/r02a_ncu.log 2>&1; echo "ncu rc=$?"
head -c 600 gpurun_out/r02a_ref.json; echo; head -c 900 gpurun_out/r02a_ours.json; echo
