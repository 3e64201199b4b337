#!/bin/bash
# One-GPU evidence run for a round: GPU tests, smoke, bench line, phase profile, ncu metric passes
# over the forward list / per-plane post-processing / consensus, full captures of the dominant
# post kernels, launch list of the bench command (256^3). ROUND_TAG names the outputs.
set -x
mkdir -p gpurun_out
R=${ROUND_TAG:-r02}
python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/tests_${R}.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_${R}.json 2> gpurun_out/bench_${R}.err
tail -c 600 gpurun_out/bench_${R}.json; tail -3 gpurun_out/bench_${R}.err
if [ "${SKIP_NCU:-0}" = "1" ]; then exit 0; fi
B200_EMPANADA_PROFILE=1 python tools/profile_pipeline.py 1024 16 > gpurun_out/phases_${R}.txt 2>&1
tail -4 gpurun_out/phases_${R}.txt
NCU="ncu --clock-control none --profile-from-start off"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active"
timeout 400 $NCU --metrics $M --csv --log-file gpurun_out/post_metrics_${R}.csv python tools/profile_post.py 1024 32 > gpurun_out/ncu_post_${R}.log 2>&1
tail -2 gpurun_out/ncu_post_${R}.log
timeout 400 $NCU --metrics $M --csv --log-file gpurun_out/cons_metrics_${R}.csv python tools/ncu_consensus.py 512 > gpurun_out/ncu_cons_${R}.log 2>&1
tail -2 gpurun_out/ncu_cons_${R}.log
timeout 500 $NCU --metrics $M --csv --log-file gpurun_out/fwd_metrics_${R}.csv python tools/ncu_forward.py 37 1024 > gpurun_out/ncu_fwd_${R}.log 2>&1
tail -2 gpurun_out/ncu_fwd_${R}.log
timeout 300 $NCU --set full --import-source on -k regex:"median|group_flags|rowruns|runs_cc|runs_paint|triple_runs" -c 10 -f -o gpurun_out/post_full_${R} python tools/profile_post.py 1024 32 > /dev/null 2>&1
BENCH_CUDA_PROFILER_API=1 timeout 600 $NCU --metrics gpu__time_duration.sum --csv \
    --log-file gpurun_out/launches_${R}.csv python bench.py --size 256 --steps 1 --warmup 1 --no-cpu --no-2d > gpurun_out/ncu_bench_${R}.log 2>&1
wc -l gpurun_out/launches_${R}.csv
ls -la gpurun_out; du -sh gpurun_out
