#!/bin/bash
# evidence run 1: one ncu --set full pass over the product's kernels in the kernel zoo (after a plain run), then the config-5 op sweep
mkdir -p gpurun_out
timeout 300 python tools/kernel_zoo.py 1 > gpurun_out/r02b_zoo_plain.log 2>&1; echo "zoo plain rc=$?"; tail -1 gpurun_out/r02b_zoo_plain.log
timeout 600 ncu --set full --clock-control none --kernel-name-base demangled -k regex:epnet:: -c 90 -o /tmp/r02b_zoo python tools/kernel_zoo.py 1 > gpurun_out/r02b_zoo_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r02b_zoo_ncu.log
ncu -i /tmp/r02b_zoo.ncu-rep --page raw --csv > gpurun_out/r02b_zoo_raw.csv 2> gpurun_out/r02b_zoo_raw.err; echo "export rc=$?"
ls -la /tmp/r02b_zoo.ncu-rep gpurun_out/r02b_zoo_raw.csv
python tools/ncu_table.py gpurun_out/r02b_zoo_raw.csv > gpurun_out/r02b_ncu_kernel_table.txt 2>&1; wc -l gpurun_out/r02b_ncu_kernel_table.txt
timeout 600 python tests/perf/op_sweep.py gpurun_out/r02b_op_sweep.json > gpurun_out/r02b_op_sweep.log 2>&1; echo "sweep rc=$?"
du -sh gpurun_out
