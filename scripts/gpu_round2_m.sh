#!/bin/bash
# pass m: fused ConvTranspose->Conv3x3 (upconv_pair_kernel): op parity, network parity, per-op times, A/B bench
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_upconv_gpu.py -m gpu -q --timeout=120 --tb=short 2>&1 | tail -25
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_real_images_gpu.py -m gpu -q -x --timeout=300 --tb=short 2>&1 | tail -12
timeout 300 python scripts/profile_ops.py 5 > gpurun_out/r2m_profile_ops.txt 2>&1; grep -E "^0/.*(up|dec)|serial" gpurun_out/r2m_profile_ops.txt
VPB_UPCONV=0 timeout 300 python scripts/profile_ops.py 5 > gpurun_out/r2m_profile_ops_unfused.txt 2>&1; grep -E "^0/.*(up|dec)|serial" gpurun_out/r2m_profile_ops_unfused.txt
for tag in fused unfused; do
  if [ $tag = unfused ]; then export VPB_UPCONV=0; fi
  timeout 600 python bench.py --steps 200 --no-cpu-baseline > gpurun_out/r2m_bench_$tag.json 2> gpurun_out/r2m_bench_$tag.err; python -c "
import json; d=json.load(open('gpurun_out/r2m_bench_$tag.json')); print('$tag', d['value'], d['e2e']['value'], d['e2e']['p50_latency_ms'], d['config']['gflop_per_frame_executed'])
for s in d['roofline']['stages']:
    if 'conv' in s['kernel']: print(s['kernel'], s['launches_per_frame'], round(s['us_per_frame'],1), round(s['achieved'],1))"
done
