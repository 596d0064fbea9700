#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_real_images_gpu.py -m gpu -q -s --timeout=300 --tb=line 2>&1 | grep -E "sigma|passed|failed|FAILED|Error" > gpurun_out/r2f_real.log
cat gpurun_out/r2f_real.log | tail -90
timeout 600 python -m pytest tests/test_autospeed_gpu.py tests/test_post_ops_gpu.py -m gpu -q --timeout=300 --tb=short -s 2>&1 | grep -E "autospeed frame|passed|failed|FAILED|Error" > gpurun_out/r2f_autospeed.log
cat gpurun_out/r2f_autospeed.log
