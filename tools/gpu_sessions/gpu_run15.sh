#!/bin/bash
# round-2 run 15: bisection stage B (tests + timing), GK-order tests, eig suite, then ncu: launch list of the bench command and
# full captures of four kernels exported to CSV on the box (the .ncu-rep files stay there: gpurun_out is capped at 64 MiB)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gk_orders.py tests/test_gpu_round2.py -x -q -k "gk_order or stage_b or bisection" > gpurun_out/r2_15_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2_15_tests.log
timeout 900 python tools/time_eig_stage_b.py > gpurun_out/r2_15_stage_b.log 2>&1; echo "stage_b rc=$?"; cat gpurun_out/r2_15_stage_b.log | tail -30
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_15_suite.log 2>&1; echo "suite rc=$?"; tail -5 gpurun_out/r2_15_suite.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_15_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-other-configs --no-check > gpurun_out/r2_15_ncu_bench.log 2>&1; echo "launches rc=$?"
mkdir -p /tmp/rep
for spec in "eig_tridiag_reg64:eig:eig" "eig_bisect:eig:bisect" "resolvent_mma_team:team:team" "iai_mid:mid:mid"; do
  IFS=: read -r rx case name <<< "$spec"
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$rx -c 1 -f -o /tmp/rep/$name python tools/profile_cases.py $case > gpurun_out/r2_15_ncu_$name.log 2>&1; echo "$name rc=$?"
  ncu -i /tmp/rep/$name.ncu-rep --page raw --csv > gpurun_out/r2_15_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/rep/$name.ncu-rep --page source --csv > gpurun_out/r2_15_${name}_source.csv 2>/dev/null
done
ls -la gpurun_out/ | head -40; du -sh gpurun_out
