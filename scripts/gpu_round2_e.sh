#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_real_images_gpu.py -m gpu -q -x --timeout=300 --tb=short 2>&1 | tail -40 > gpurun_out/r2e_real.log
timeout 600 python -m pytest tests/test_real_images_gpu.py -m gpu -q --timeout=300 --tb=line 2>&1 | tail -15 >> gpurun_out/r2e_real.log
cat gpurun_out/r2e_real.log | tail -30
timeout 900 python -m pytest tests/test_autospeed_gpu.py -m gpu -q --timeout=300 --tb=short -s 2>&1 | tail -80 > gpurun_out/r2e_autospeed.log
tail -60 gpurun_out/r2e_autospeed.log
