#!/bin/bash
python tools/gpu_probe.py > gpurun_out/r2_probe.log 2>&1
for v in 0 1; do
  ABZ_MMA_VARIANT=$v python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs --no-check > gpurun_out/r2_b_var$v.json 2> gpurun_out/r2_b_var$v.err
done
ABZ_MMA_VARIANT=1 ABZ_MMA_WARPS=12 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs --no-check > gpurun_out/r2_b_var1_w12.json 2> gpurun_out/r2_b_var1_w12.err
cat gpurun_out/r2_probe.log
for f in gpurun_out/r2_b_var0.json gpurun_out/r2_b_var1.json gpurun_out/r2_b_var1_w12.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['value'], d['roofline']['frac'], d['roofline']['matfun_ms_per_step'])"; done
