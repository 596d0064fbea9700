#!/bin/bash
# ncu --set full of one conv layer: bash scripts/ncu_one.sh <tag> <run_layer.py args...>
mkdir -p gpurun_out
tag=$1; shift
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'conv' -s 2 -c 1 -f \
   -o gpurun_out/ncu4_$tag python scripts/run_layer.py "$@" > gpurun_out/ncu4_$tag.log 2>&1
tail -1 gpurun_out/ncu4_$tag.log
