#!/bin/bash
python tools/probe_team_accuracy.py > gpurun_out/r2_team_acc_default.log 2>&1
ABZ_MMA_TEAM_PIVOT_THR=0 python tools/probe_team_accuracy.py > gpurun_out/r2_team_acc_thr0.log 2>&1
python tools/profile_cases.py eig > gpurun_out/r2_eig_time.log 2>&1
python tools/profile_cases.py eig >> gpurun_out/r2_eig_time.log 2>&1
python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r2_t5.log
cat gpurun_out/r2_team_acc_default.log gpurun_out/r2_team_acc_thr0.log gpurun_out/r2_eig_time.log gpurun_out/r2_t5.log
