#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/time_c3.py > gpurun_out/r2_17_c3.log 2>&1; echo "c3 rc=$?"; cat gpurun_out/r2_17_c3.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_17_suite.log 2>&1; echo "suite rc=$?"; tail -5 gpurun_out/r2_17_suite.log
