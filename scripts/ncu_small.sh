#!/bin/bash
# ncu --set full on the small (non-GEMM) kernels of one SceneSeg frame
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'preprocess|stem|depthwise|se_scale|gap_kernel' -s 20 -c 12 -f \
   -o gpurun_out/ncu_small python scripts/dev_e2e.py scene_seg > gpurun_out/ncu_small.log 2>&1
tail -3 gpurun_out/ncu_small.log
