#!/bin/bash
# pass n: network parity with the fused upconv layers + in-kernel timelines of the fused layers
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_real_images_gpu.py tests/test_split_precision_gpu.py -m gpu -q --timeout=300 --tb=short 2>&1 | tail -12
for cfg in "160 320 128 128 0" "160 320 128 128 0 64" "80 160 256 256 32" "80 160 256 256 32 128" "40 80 512 512 24" "40 80 512 512 24 128" "20 40 768 512 40" "20 40 768 512 40 128" "10 20 1280 768 80" "10 20 1280 768 80 128" "10 20 1280 768 80 64"; do
  echo "== $cfg"; timeout 120 python scripts/trace_upconv.py $cfg
done > gpurun_out/r2n_trace_upconv.txt 2>&1
cat gpurun_out/r2n_trace_upconv.txt
