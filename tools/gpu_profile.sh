#!/bin/bash
# Round evidence run (one B200): GPU tests, bench line, ncu launch list of the same bench command,
# and `--set full` captures of the forward launch list, the per-plane post kernels and consensus.
set -x
mkdir -p gpurun_out
R=${ROUND_TAG:-r01}
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_${R}.json 2> gpurun_out/bench_${R}.err
tail -c 3000 gpurun_out/bench_${R}.json
B200_EMPANADA_PROFILE=1 python tools/profile_pipeline.py 1024 16 > gpurun_out/phases_${R}.txt 2>&1
tail -4 gpurun_out/phases_${R}.txt
python tools/profile_forward.py 16 1024 > gpurun_out/prof_fwd_${R}.txt 2>&1
head -12 gpurun_out/prof_fwd_${R}.txt
NCU="ncu --clock-control none --profile-from-start off"
BENCH_CUDA_PROFILER_API=1 timeout 900 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_${R}.csv \
    python bench.py --size 512 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_bench_${R}.log 2>&1
timeout 600 $NCU --set full --import-source on -f -o gpurun_out/fwd_full_${R} python tools/ncu_forward.py 16 1024 > gpurun_out/ncu_fwd_${R}.log 2>&1
timeout 600 $NCU --set full --import-source on -f -o gpurun_out/post_full_${R} python tools/profile_post.py 1024 32 > gpurun_out/ncu_post_${R}.log 2>&1
timeout 600 $NCU --set full --import-source on -f -o gpurun_out/cons_full_${R} python tools/ncu_consensus.py 512 > gpurun_out/ncu_cons_${R}.log 2>&1
ls -la gpurun_out
