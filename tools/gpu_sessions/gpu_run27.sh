#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused_mma.py tests/test_gpu_multitile.py -m gpu -x -q > gpurun_out/r2_27_tests_fused.log 2>&1; echo "fused tests rc=$?"; tail -n 3 gpurun_out/r2_27_tests_fused.log
ABZ_MMA_WARPS=12 timeout 600 python -m pytest tests/test_gpu_fused_mma.py -m gpu -x -q > gpurun_out/r2_27_tests_fused_w12.log 2>&1; echo "fused w12 tests rc=$?"; tail -n 3 gpurun_out/r2_27_tests_fused_w12.log
for w in 8 12; do
  ABZ_MMA_WARPS=$w timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/r2_27_b_w$w.json 2> gpurun_out/r2_27_b_w$w.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_27_b_w$w.json"))
print("WARPS $w", d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["eval_ms_per_step"], d["roofline"]["matfun_ms_per_step"], d["check"]["rel_err_vs_cpu"])
PY
done
