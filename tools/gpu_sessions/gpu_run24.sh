#!/bin/bash
mkdir -p gpurun_out
ABZ_MMA_VARIANT=4 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_24_tests_var4.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2_24_tests_var4.log
for v in 4; do
  ABZ_MMA_VARIANT=$v timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/r2_24_b_var$v.json 2> gpurun_out/r2_24_b_var$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_24_b_var$v.json"))
print("VAR $v", d["value"], d["roofline"]["frac"], d["roofline"]["matfun_ms_per_step"], d["check"]["rel_err_vs_cpu"])
PY
done
