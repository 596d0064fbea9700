#!/bin/bash
# fourth GPU pass (1 GPU): split-mode accuracy printout, bench, ncu evidence exported to CSV on the box (the .ncu-rep
# files of 80 kernels exceed the 64 MiB return limit)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_split_precision_gpu.py tests/test_real_images_gpu.py -m gpu -q -s --timeout=300 2>&1 | grep -E "sigma|passed|failed|FAILED" > gpurun_out/r2d_split_real.log
tail -12 gpurun_out/r2d_split_real.log
timeout 600 python bench.py --steps 200 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"
exp() { ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/$1.csv 2>/dev/null; ls -la /tmp/$1.ncu-rep | awk '{print $5}'; }
timeout 900 ncu --set full --clock-control none \
   -k regex:'preprocess|stem|depthwise|se_scale|gap_kernel|linear_kernel|ctx_conv1|fuse_pool' -s 81 -c 82 -f \
   -o /tmp/r2_ncu_hbm_stages python scripts/profile_ops.py 1 > gpurun_out/r2_ncu_hbm_stages.log 2>&1; echo "ncu hbm rc=$?"; exp r2_ncu_hbm_stages
timeout 300 ncu --set full --clock-control none -k regex:'conv_gemm' -s 2 -c 1 -f \
   -o /tmp/r2_ncu_up3 python scripts/run_layer.py up3 80 160 256 256 1 4 > gpurun_out/r2_ncu_up3.log 2>&1; echo "ncu up3 rc=$?"; exp r2_ncu_up3
timeout 300 ncu --set full --clock-control none -k regex:'conv_gemm' -s 2 -c 1 -f \
   -o /tmp/r2_ncu_up4 python scripts/run_layer.py up4 160 320 128 128 1 4 > gpurun_out/r2_ncu_up4.log 2>&1; echo "ncu up4 rc=$?"; exp r2_ncu_up4
timeout 300 ncu --set full --clock-control none -k regex:'lateral|visualize' -s 4 -c 4 -f \
   -o /tmp/r2_ncu_post python scripts/bench_post.py > gpurun_out/r2_ncu_post.log 2>&1; echo "ncu post rc=$?"; exp r2_ncu_post
bash scripts/ncu_launch_list.sh > gpurun_out/r2_launch_list.txt 2>&1; tail -20 gpurun_out/r2_launch_list.txt | head -16
bash scripts/ncu_conv_traffic.sh > gpurun_out/r2_conv_traffic.txt 2>&1; tail -2 gpurun_out/r2_conv_traffic.txt
du -sh gpurun_out
