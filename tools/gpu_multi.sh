#!/bin/bash
# N-GPU validation: parity of DistributedEngine3d vs single GPU, then the bench at N ranks.
set -x
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
python -m pytest tests/test_gpu_multigpu.py -m gpu -x -q -s 2>&1 | tail -15
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 \
    bench.py --gpus $N --steps 2 --warmup 2 --no-cpu --no-2d > gpurun_out/bench_n${N}.json 2> gpurun_out/bench_n${N}.err
tail -c 1500 gpurun_out/bench_n${N}.json; tail -5 gpurun_out/bench_n${N}.err
