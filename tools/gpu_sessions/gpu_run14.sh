#!/bin/bash
# round-2 run 14: GK-order tests, ncu launch list of the bench command, full captures of the eig / team / mid kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gk_orders.py tests/test_gpu_matrix_iai.py -x -q > gpurun_out/r2_14_gk.log 2>&1; echo "gk rc=$?"; tail -5 gpurun_out/r2_14_gk.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_14_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-other-configs --no-check > gpurun_out/r2_14_ncu_bench.log 2>&1; echo "launches rc=$?"
for spec in "eig_tridiag_reg64:eig:eig" "eig_tql:eig:tql" "resolvent_mma_team:team:team" "iai_mid:mid:mid"; do
  IFS=: read -r rx case name <<< "$spec"
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$rx -c 1 -f -o gpurun_out/r2_14_$name python tools/profile_cases.py $case > gpurun_out/r2_14_ncu_$name.log 2>&1; echo "$name rc=$?"
done
ls -la gpurun_out/*.ncu-rep
