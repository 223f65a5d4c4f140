#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused_mma.py tests/test_gpu_multitile.py -m gpu -x -q > gpurun_out/r2_26_tests_fused.log 2>&1; echo "fused tests rc=$?"; tail -n 3 gpurun_out/r2_26_tests_fused.log
for f in 1; do
  ABZ_FUSED_MMA=$f timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/r2_26_b_fused$f.json 2> gpurun_out/r2_26_b_fused$f.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_26_b_fused$f.json"))
print("FUSED $f", d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["eval_ms_per_step"], d["roofline"]["matfun_ms_per_step"], d["check"]["rel_err_vs_cpu"])
PY
done
