#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_iai_middles.py tests/test_gpu_parity.py -x -q -k "iai or lookahead or middles" > gpurun_out/r2_18_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r2_18_tests.log
timeout 600 python tools/time_c3.py > gpurun_out/r2_18_c3.log 2>&1; echo "c3 rc=$?"; cat gpurun_out/r2_18_c3.log
