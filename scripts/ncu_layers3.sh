#!/bin/bash
# ncu --set full of the round-1 final conv kernels (pair kernel, ConvT tile kernel) -> gpurun_out/ncu3_*.ncu-rep
mkdir -p gpurun_out
run() {
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:'conv' -s 2 -c 1 -f \
     -o gpurun_out/ncu3_$1 python scripts/run_layer.py "$@" > gpurun_out/ncu3_$1.log 2>&1
  tail -1 gpurun_out/ncu3_$1.log
}
run dec6 160 320 256 256 9 1 0
run dec8 320 640 128 128 9 1 0
run up3 80 160 256 256 1 4 0
run up2 40 80 512 512 1 4 0
run dec0 20 40 1280 768 9 1 0
run dec10 320 640 64 3 9 1 0
