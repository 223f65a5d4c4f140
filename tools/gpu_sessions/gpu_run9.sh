#!/bin/bash
timeout 600 ncu --set full --import-source on --clock-control none -k regex:resolvent_mma -c 1 -o gpurun_out/r2_k3fast python tools/profile_cases.py contract > gpurun_out/r2_ncu_k3fast.log 2>&1
ls -la gpurun_out/r2_k3fast.ncu-rep
ncu -i gpurun_out/r2_k3fast.ncu-rep --page source --csv --print-source sass > gpurun_out/r2_k3fast_source.csv 2>/dev/null
ncu -i gpurun_out/r2_k3fast.ncu-rep --page raw --csv > gpurun_out/r2_k3fast_raw.csv 2>/dev/null
wc -l gpurun_out/r2_k3fast_source.csv gpurun_out/r2_k3fast_raw.csv
