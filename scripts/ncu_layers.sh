#!/bin/bash
# ncu --set full captures of representative conv layers (run under gpurun; writes gpurun_out/ncu_*.ncu-rep)
mkdir -p gpurun_out
run() { # name H W Cin Cout taps phases mode
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 2 -c 1 -f \
     -o gpurun_out/ncu_$1 python scripts/run_layer.py "$@" > gpurun_out/ncu_$1.log 2>&1
  tail -2 gpurun_out/ncu_$1.log
}
run dec8 320 640 128 128 9 1 0
run up4 160 320 128 128 1 4 0
run skip3 160 320 32 256 1 1 0
run dec6 160 320 256 256 9 1 0
run dec0 20 40 1280 768 9 1 0
