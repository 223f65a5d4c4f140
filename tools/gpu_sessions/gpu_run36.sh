#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py tests/test_gpu_fused_mma.py tests/test_gpu_dos.py -m gpu -x -q > gpurun_out/r2_36_tests.log 2>&1; echo "tests rc=$?"; tail -n 8 gpurun_out/r2_36_tests.log
