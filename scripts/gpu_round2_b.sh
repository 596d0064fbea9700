#!/bin/bash
# second GPU pass (2 GPUs): parity suite (new preprocess + TMA-store ConvT + config 5), conv micro-bench A/B, op profile, bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -8 gpurun_out/r2b_pytest.log
timeout 300 python scripts/bench_conv.py lin > gpurun_out/r2b_bench_conv.txt 2>&1; tail -1 gpurun_out/r2b_bench_conv.txt
timeout 300 python scripts/profile_ops.py 5 > gpurun_out/r2b_profile_ops.txt 2>&1; head -2 gpurun_out/r2b_profile_ops.txt
timeout 600 python bench.py --steps 200 --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
