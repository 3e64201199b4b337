#!/bin/bash
set -x
mkdir -p gpurun_out
R=${ROUND_TAG:-r01}
NCU="ncu --clock-control none --profile-from-start off"
timeout 300 $NCU --set full --import-source on -k regex:"dwconv|stem_pool" -c 4 -f -o gpurun_out/dw_stem2_${R} python tools/ncu_forward.py 16 1024 > /dev/null 2>&1
ls -la gpurun_out
