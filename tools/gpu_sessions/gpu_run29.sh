#!/bin/bash
# K3-fused with the direct (materialised) mode: tests, ncu full capture of one fused launch, launch list of the bench command, bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_29_tests_all.log 2>&1; echo "all tests rc=$?"; tail -n 3 gpurun_out/r2_29_tests_all.log
python tools/profile_cases.py fused > gpurun_out/r2_29_fused_plain.log 2>&1 && \
timeout 600 ncu --set full --import-source on --clock-control none -k regex:resolvent_mma_fused -c 1 -o gpurun_out/r2_29_fused -f python tools/profile_cases.py fused > gpurun_out/r2_29_fused_ncu.log 2>&1; echo "ncu rc=$?"
timeout 300 python bench.py --steps 2 --warmup 3 > gpurun_out/r2_29_bench.json 2> gpurun_out/r2_29_bench.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_29_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-other-configs > gpurun_out/r2_29_ncu_bench.log 2>&1; echo "launch list rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2_29_bench.json"))
print(d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["eval_ms_per_step"], d["roofline"]["matfun_ms_per_step"], d["check"]["rel_err_vs_cpu"])
PY
