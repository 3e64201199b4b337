#!/bin/bash
set -x
N=${1:-4}
mkdir -p gpurun_out
B200_EMPANADA_PROFILE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29546 \
    bench.py --gpus $N --steps 1 --warmup 1 --no-cpu --no-2d > gpurun_out/prof_n${N}.txt 2> gpurun_out/prof_n${N}.err
grep "^\[rank" gpurun_out/prof_n${N}.txt | tail -40; grep '"metric"' gpurun_out/prof_n${N}.txt | cut -c1-300
