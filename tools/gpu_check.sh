#!/bin/bash
# Quick GPU pass: tests, per-op forward profile, bench line, launch list of the bench command.
set -x
mkdir -p gpurun_out
R=${ROUND_TAG:-r01}
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python tools/profile_forward.py 16 1024 > gpurun_out/prof_fwd_${R}.txt 2>&1
head -14 gpurun_out/prof_fwd_${R}.txt
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_${R}.json 2> gpurun_out/bench_${R}.err
tail -c 3500 gpurun_out/bench_${R}.json; tail -5 gpurun_out/bench_${R}.err
B200_EMPANADA_PROFILE=1 python tools/profile_pipeline.py 1024 16 > gpurun_out/phases_${R}.txt 2>&1
tail -6 gpurun_out/phases_${R}.txt
BENCH_CUDA_PROFILER_API=1 timeout 600 ncu --clock-control none --profile-from-start off --metrics gpu__time_duration.sum --csv \
    --log-file gpurun_out/launches_${R}.csv python bench.py --size 256 --steps 1 --warmup 1 --no-cpu --no-2d > gpurun_out/ncu_bench_${R}.log 2>&1
tail -2 gpurun_out/ncu_bench_${R}.log
ls -la gpurun_out; du -sh gpurun_out
