#!/bin/bash
python -m pytest tests/test_gpu_general_limits.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_t4.log
python -m pytest tests -m gpu -x -q 2>&1 | tail -12 >> gpurun_out/r2_t4.log
cat gpurun_out/r2_t4.log
