#!/bin/bash
# Bench at N ranks only (no pytest).
set -x
N=${1:-4}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 \
    bench.py --gpus $N --steps 2 --warmup 2 --no-cpu --no-2d > gpurun_out/bench_n${N}.json 2> gpurun_out/bench_n${N}.err
grep '"metric"' gpurun_out/bench_n${N}.json | tail -c 1200; tail -3 gpurun_out/bench_n${N}.err
