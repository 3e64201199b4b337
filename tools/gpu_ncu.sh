#!/bin/bash
# ncu evidence (one B200): cheap metric passes over every launch of the forward list / the post
# kernels / consensus (DRAM bytes, duration, tensor-pipe and DRAM utilisation), plus `--set full`
# captures of a handful of representative launches. Keeps gpurun_out small.
set -x
mkdir -p gpurun_out
R=${ROUND_TAG:-r01}
NCU="ncu --clock-control none --profile-from-start off"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active"
timeout 500 $NCU --metrics $M --csv --log-file gpurun_out/fwd_metrics_${R}.csv python tools/ncu_forward.py 37 1024 > gpurun_out/ncu_fwd_${R}.log 2>&1
tail -3 gpurun_out/ncu_fwd_${R}.log
timeout 400 $NCU --metrics $M --csv --log-file gpurun_out/post_metrics_${R}.csv python tools/profile_post.py 1024 32 > gpurun_out/ncu_post_${R}.log 2>&1
tail -2 gpurun_out/ncu_post_${R}.log
timeout 400 $NCU --metrics $M --csv --log-file gpurun_out/cons_metrics_${R}.csv python tools/ncu_consensus.py 512 > gpurun_out/ncu_cons_${R}.log 2>&1
tail -2 gpurun_out/ncu_cons_${R}.log
# full captures: ASPP 3x3 (deep K), a layer-1 residual 1x1, the fused-head pointwise, depthwise, stem
timeout 300 $NCU --set full -k regex:conv_gemm -s 54 -c 2 -f -o gpurun_out/conv_aspp_${R} python tools/ncu_forward.py 37 1024 > /dev/null 2>&1
timeout 300 $NCU --set full -k regex:conv_gemm -s 3 -c 2 -f -o gpurun_out/conv_layer1_${R} python tools/ncu_forward.py 37 1024 > /dev/null 2>&1
timeout 300 $NCU --set full -k regex:"dwconv|stem_pool" -c 3 -f -o gpurun_out/dw_stem_${R} python tools/ncu_forward.py 37 1024 > /dev/null 2>&1
timeout 300 $NCU --set full -k regex:"median|group_pixels|cc_|pair_overlap|relabel|runs_" -c 14 -f -o gpurun_out/post_full_${R} python tools/profile_post.py 1024 32 > /dev/null 2>&1
ls -la gpurun_out; du -sh gpurun_out
