#!/bin/bash
# GPU session script: tests, a short bench, compute-sanitizer attempts on the multi-tile cases (results under gpurun_out/)
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_t1.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_b1.json 2> gpurun_out/r2_b1.err
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_cases.py multitile > gpurun_out/r2_sanitize_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r2_sanitize_racecheck.log
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_cases.py multitile > gpurun_out/r2_sanitize_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r2_sanitize_memcheck.log
tail -3 gpurun_out/r2_t1.log
