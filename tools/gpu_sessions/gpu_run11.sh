#!/bin/bash
run() { env "$@" timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs --no-check 2> gpurun_out/r2_b11.err | python -c "
import json,sys
t=sys.stdin.read().strip().splitlines()
if not t: print('$* FAILED'); sys.exit(0)
d=json.loads(t[-1]); print('$*', d['value'], d['roofline']['frac'], d['roofline']['matfun_ms_per_step'])"; tail -3 gpurun_out/r2_b11.err; }
run ABZ_MMA_VARIANT=1 ABZ_MMA_WARPS=4
run ABZ_MMA_VARIANT=3 ABZ_MMA_WARPS=4
