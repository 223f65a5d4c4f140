#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -x -q -k "eig or stage_b or bisection or sweep or tridiag" > gpurun_out/r2_22_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2_22_tests.log
for sp in 0 1; do echo "ABZ_TRIDIAG_SPLIT=$sp"; ABZ_TRIDIAG_SPLIT=$sp timeout 600 python tools/time_eig_stage_b.py 2>&1 | grep -E "n=64|n=48"; done > gpurun_out/r2_22_split.log 2>&1; cat gpurun_out/r2_22_split.log
