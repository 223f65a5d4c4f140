#!/bin/bash
# final-state validation of the K3-fused build: smoke, whole GPU suite, full bench line (cpu baseline + other configs), reference arm
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_35_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/r2_35_smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_35_tests_all.log 2>&1; echo "all tests rc=$?"; tail -n 3 gpurun_out/r2_35_tests_all.log
timeout 600 python bench.py > gpurun_out/r2_35_bench.json 2> gpurun_out/r2_35_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_35_bench_ref.json 2> gpurun_out/r2_35_bench_ref.err; echo "ref rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2_35_bench.json"))
print(d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["eval_ms_per_step"], d["roofline"]["matfun_ms_per_step"], d["check"]["rel_err_vs_cpu"], d["gpu_launches"])
for o in d["other_configs"]: print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in o.items() if k != "ms_all_repetitions"})
r=json.load(open("gpurun_out/r2_35_bench_ref.json")); print("ref", r["value"], r["cpu_baseline"]["cores"])
PY
