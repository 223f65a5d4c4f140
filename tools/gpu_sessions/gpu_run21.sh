#!/bin/bash
mkdir -p gpurun_out
( echo "# round 2: python tools/fuzz_parity.py <cases> <seed> / tools/fuzz_iai.py <cases> <seed> on one B200"
  timeout 900 python tools/fuzz_parity.py 700 23 2>&1 | tail -4
  timeout 900 python tools/fuzz_iai.py 300 21 2>&1 | tail -4
  timeout 900 python tools/fuzz_iai.py 150 22 2>&1 | tail -4 ) > gpurun_out/r2_21_fuzz.log 2>&1
cat gpurun_out/r2_21_fuzz.log
