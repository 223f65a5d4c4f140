#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused_mma.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_30_tests.log 2>&1; echo "tests rc=$?"; tail -n 2 gpurun_out/r2_30_tests.log
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/r2_30_bench.json 2> gpurun_out/r2_30_bench.err; echo "bench rc=$?"
python tools/profile_cases.py fused > gpurun_out/r2_30_fused_plain.log 2>&1 && \
timeout 600 ncu --set full --import-source on --clock-control none -k regex:resolvent_mma_fused -c 1 -o gpurun_out/r2_30_fused -f python tools/profile_cases.py fused > gpurun_out/r2_30_fused_ncu.log 2>&1; echo "ncu rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2_30_bench.json"))
print(d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["eval_ms_per_step"], d["roofline"]["matfun_ms_per_step"], d["check"]["rel_err_vs_cpu"])
PY
cat gpurun_out/r2_30_fused_plain.log
