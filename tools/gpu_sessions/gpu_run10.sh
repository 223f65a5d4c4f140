#!/bin/bash
run() { env "$@" timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs --no-check 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', d['value'], d['roofline']['frac'], d['roofline']['matfun_ms_per_step'])"; }
run ABZ_MMA_VARIANT=1 ABZ_MMA_WARPS=4
run ABZ_MMA_VARIANT=1 ABZ_MMA_WARPS=8
run ABZ_MMA_VARIANT=3 ABZ_MMA_WARPS=8
run ABZ_MMA_VARIANT=3 ABZ_MMA_WARPS=12
ABZ_MMA_VARIANT=3 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -m gpu -q 2>&1 | tail -4
