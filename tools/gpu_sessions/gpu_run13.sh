#!/bin/bash
# round-2 run 13 (N GPUs): the bench line at N ranks (allreduce inside the timed loop, other_configs collectively) + multi_gpu_check
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_13_bench_n$N.json 2> gpurun_out/r2_13_bench_n$N.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/r2_13_bench_n$N.json; tail -5 gpurun_out/r2_13_bench_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 tools/multi_gpu_check.py > gpurun_out/r2_13_check_n$N.log 2> gpurun_out/r2_13_check_n$N.err; echo "check rc=$?"; cat gpurun_out/r2_13_check_n$N.log; tail -5 gpurun_out/r2_13_check_n$N.err
