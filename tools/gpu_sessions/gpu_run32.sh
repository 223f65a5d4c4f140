#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_iai_norb6.py -m gpu -x -q > gpurun_out/r2_32_tests_norb6.log 2>&1; echo "norb6 tests rc=$?"; tail -n 15 gpurun_out/r2_32_tests_norb6.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_iai_middles.py tests/test_gpu_general_limits.py tests/test_gpu_matrix_iai.py tests/test_gpu_gk_orders.py -m gpu -x -q > gpurun_out/r2_32_tests_iai.log 2>&1; echo "iai tests rc=$?"; tail -n 3 gpurun_out/r2_32_tests_iai.log
timeout 600 python tools/time_iai_norb.py 0.02 1e-5 > gpurun_out/r2_32_norb6_timing.log 2>&1
cat gpurun_out/r2_32_norb6_timing.log
