#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gemm_gpu.py -m gpu -q --timeout=300 --tb=short 2>&1 | tail -6
exp() { ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/$1.csv 2>/dev/null; }
timeout 300 ncu --set full --clock-control none -k regex:'convt_ws' -s 2 -c 1 -f -o /tmp/r2_ncu_ws_up3 python scripts/run_layer.py up3 80 160 256 256 1 4 > gpurun_out/r2_ncu_ws_up3.log 2>&1; exp r2_ncu_ws_up3
timeout 300 ncu --set full --clock-control none -k regex:'convt_ws' -s 2 -c 1 -f -o /tmp/r2_ncu_ws_up4 python scripts/run_layer.py up4 160 320 128 128 1 4 > gpurun_out/r2_ncu_ws_up4.log 2>&1; exp r2_ncu_ws_up4
timeout 300 python scripts/bench_conv.py lin > gpurun_out/r2j_bench_conv.txt 2>&1; grep -E "up[0-9]" gpurun_out/r2j_bench_conv.txt
