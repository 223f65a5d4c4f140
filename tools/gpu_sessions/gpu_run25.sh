#!/bin/bash
# K3-fused (stage 1 folded into the DMMA resolvent kernel): new tests, whole GPU suite, bench fused vs unfused
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused_mma.py -m gpu -x -q > gpurun_out/r2_25_tests_fused.log 2>&1; echo "fused tests rc=$?"; tail -n 12 gpurun_out/r2_25_tests_fused.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_25_tests_all.log 2>&1; echo "all tests rc=$?"; tail -n 4 gpurun_out/r2_25_tests_all.log
for f in 1 0; do
  ABZ_FUSED_MMA=$f timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/r2_25_b_fused$f.json 2> gpurun_out/r2_25_b_fused$f.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_25_b_fused$f.json"))
print("FUSED $f", d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["eval_ms_per_step"], d["roofline"]["matfun_ms_per_step"], d["check"]["rel_err_vs_cpu"])
PY
done
