#!/bin/bash
# N GPUs of one box: sharded parity check, then the bench line at N (other configs run collectively: strong scaling)
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py > gpurun_out/r2_37_multi$N.log 2> gpurun_out/r2_37_multi$N.err; echo "multi check rc=$?"; tail -n 12 gpurun_out/r2_37_multi$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/r2_37_bench_n$N.json 2> gpurun_out/r2_37_bench_n$N.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2_37_bench_n$N.json") if l.startswith("{")][-1])
print(d["n_gpus"], d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["check"]["rel_err_vs_cpu"])
for o in d["other_configs"]: print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in o.items() if k not in ("ms_all_repetitions",)})
PY
