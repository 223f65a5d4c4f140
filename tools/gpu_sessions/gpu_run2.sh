#!/bin/bash
# GPU session: DMMA team resolvent - parity tests, timings against the pivoted teams, two-warp teams at norb = 32
python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "team" 2>&1 | tail -15 > gpurun_out/r2_t2.log
python tools/time_team_resolvent.py 64 56 48 40 > gpurun_out/r2_team_timing.log 2>&1
python tools/time_team_resolvent.py 32 >> gpurun_out/r2_team_timing.log 2>&1
ABZ_MMA_TEAM=1 python tools/time_team_resolvent.py 32 >> gpurun_out/r2_team_timing.log 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 >> gpurun_out/r2_t2.log
cat gpurun_out/r2_t2.log gpurun_out/r2_team_timing.log
