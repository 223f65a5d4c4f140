#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_iai_middles.py tests/test_gpu_round2.py -m gpu -x -q -k "middles or c3" 2>&1 | tail -15 > gpurun_out/r2_t7.log
timeout 300 python tools/time_c3.py > gpurun_out/r2_c3_timing.log 2>&1
cat gpurun_out/r2_t7.log gpurun_out/r2_c3_timing.log
