#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_20_suite.log 2>&1; echo "suite rc=$?"; tail -5 gpurun_out/r2_20_suite.log
timeout 600 python tools/time_c3.py > gpurun_out/r2_20_c3.log 2>&1; echo "c3 rc=$?"; grep '"device_middles": true' gpurun_out/r2_20_c3.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2_20_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_20_smoke.log
timeout 600 python bench.py > gpurun_out/r2_20_bench.json 2> gpurun_out/r2_20_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_20_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['check']['rel_err_vs_cpu'])
for c in d['other_configs']: print(c)
PY
