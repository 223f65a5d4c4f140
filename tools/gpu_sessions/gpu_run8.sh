#!/bin/bash
for v in 1 2; do
  ABZ_MMA_VARIANT=$v timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/r2_b8_var$v.json 2> gpurun_out/r2_b8_var$v.err
done
ABZ_MMA_VARIANT=2 ABZ_MMA_WARPS=12 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/r2_b8_var2_w12.json 2> gpurun_out/r2_b8_var2_w12.err
ABZ_MMA_VARIANT=2 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multitile.py tests/test_gpu_edge_cases.py -m gpu -q 2>&1 | tail -6 > gpurun_out/r2_t8.log
for f in gpurun_out/r2_b8_var1.json gpurun_out/r2_b8_var2.json gpurun_out/r2_b8_var2_w12.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['value'], d['roofline']['frac'], d['roofline']['matfun_ms_per_step'], d['check']['rel_err_vs_cpu'])"; done
cat gpurun_out/r2_t8.log; tail -3 gpurun_out/r2_b8_var2.err
