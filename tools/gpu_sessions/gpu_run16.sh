#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -x -q -k "stage_b or bisection or eig" > gpurun_out/r2_16_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r2_16_tests.log
timeout 900 python tools/time_eig_stage_b.py > gpurun_out/r2_16_stage_b.log 2>&1; echo "stage_b rc=$?"; cat gpurun_out/r2_16_stage_b.log | tail -30
