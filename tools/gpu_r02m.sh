#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/first_conv_probe.py > gpurun_out/r02m_probe.txt 2>&1; cat gpurun_out/r02m_probe.txt
timeout 300 ncu --set full --clock-control none --kernel-name-base demangled -k regex:first_conv -c 1 -o /tmp/r02m_fc python tools/first_conv_probe.py > gpurun_out/r02m_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/r02m_fc.ncu-rep --page raw --csv > gpurun_out/r02m_fc_raw.csv 2>/dev/null
ncu -i /tmp/r02m_fc.ncu-rep --page details > gpurun_out/r02m_fc_details.txt 2>/dev/null
ls -la gpurun_out/r02m_fc_raw.csv gpurun_out/r02m_fc_details.txt
