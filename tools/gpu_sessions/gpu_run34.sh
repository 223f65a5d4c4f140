#!/bin/bash
mkdir -p gpurun_out
ABZ_MMA_PREFETCH=3 timeout 600 python -m pytest tests/test_gpu_fused_mma.py tests/test_gpu_multitile.py -m gpu -x -q > gpurun_out/r2_34_tests_pf3.log 2>&1; echo "pf3 tests rc=$?"; tail -n 2 gpurun_out/r2_34_tests_pf3.log
for pf in 3; do
  ABZ_MMA_PREFETCH=$pf timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/r2_34_b_pf$pf.json 2> gpurun_out/r2_34_b_pf$pf.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_34_b_pf$pf.json"))
print("PF $pf", d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["eval_ms_per_step"], d["roofline"]["matfun_ms_per_step"], d["check"]["rel_err_vs_cpu"])
PY
done
