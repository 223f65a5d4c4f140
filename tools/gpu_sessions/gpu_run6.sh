#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_iai_middles.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_t6.log
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -12 >> gpurun_out/r2_t6.log
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-check > gpurun_out/r2_b6.json 2> gpurun_out/r2_b6.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_ncu_eig_launches.csv python tools/profile_cases.py eig > gpurun_out/r2_ncu_eig.log 2>&1
cat gpurun_out/r2_t6.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_b6.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline']['frac'])
for o in d['other_configs']: print(o)
PY
grep -v "^==" gpurun_out/r2_ncu_eig_launches.csv | awk -F'","' '{print $5, $NF}' | tail -20
