#!/bin/bash
# third GPU pass (1 GPU): parity suite incl. split-fp16 mode, bench, ncu evidence for the HBM-bound stages + ConvT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=300 --durations=12 > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
grep -E "passed|failed|FAILED" gpurun_out/r2c_pytest.log | tail -12
timeout 600 python bench.py --steps 200 --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"
# ncu --set full: every non-GEMM kernel of one 4-task frame (second eager pass of profile_ops.py)
timeout 900 ncu --set full --clock-control none --import-source on \
   -k regex:'preprocess|stem|depthwise|se_scale|gap_kernel|linear_kernel|ctx_conv1|fuse_pool' -s 81 -c 82 -f \
   -o gpurun_out/r2_ncu_hbm_stages python scripts/profile_ops.py 1 > gpurun_out/r2_ncu_hbm_stages.log 2>&1; echo "ncu hbm rc=$?"
# ConvTranspose + skip (TMA-store epilogue) and the output-side / lateral kernels
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'conv_gemm' -s 2 -c 1 -f \
   -o gpurun_out/r2_ncu_up3 python scripts/run_layer.py up3 80 160 256 256 1 4 > gpurun_out/r2_ncu_up3.log 2>&1; echo "ncu up3 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'lateral|visualize' -s 4 -c 4 -f \
   -o gpurun_out/r2_ncu_post python scripts/bench_post.py > gpurun_out/r2_ncu_post.log 2>&1; echo "ncu post rc=$?"
