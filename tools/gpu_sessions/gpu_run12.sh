#!/bin/bash
# round-2 run 12: matrix-valued IAI tests, true one-warp-per-SMSP measurement of K3-fast, full GPU suite, N=1 bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_matrix_iai.py -x -q > gpurun_out/r2_12_matrix.log 2>&1; echo "matrix rc=$?"; tail -5 gpurun_out/r2_12_matrix.log
bash tools/gpu_run11.sh > gpurun_out/r2_12_w4.log 2>&1; cat gpurun_out/r2_12_w4.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_12_suite.log 2>&1; echo "suite rc=$?"; tail -5 gpurun_out/r2_12_suite.log
timeout 600 python bench.py > gpurun_out/r2_12_bench.json 2> gpurun_out/r2_12_bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/r2_12_bench.json
