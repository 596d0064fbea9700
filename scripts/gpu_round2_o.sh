#!/bin/bash
# pass o: upconv direct-store epilogue (deeper ring) vs staged TMA store
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_upconv_gpu.py -m gpu -q --timeout=120 --tb=short 2>&1 | tail -8
for cfg in "160 320 128 128 0 0" "160 320 128 128 0 0 3" "80 160 256 256 32 0" "80 160 256 256 32 0 3" "40 80 512 512 24 0" "40 80 512 512 24 0 3" "20 40 768 512 40 0" "20 40 768 512 40 128" "10 20 1280 768 80 0" "10 20 1280 768 80 128"; do
  echo "== $cfg"; timeout 120 python scripts/trace_upconv.py $cfg
done > gpurun_out/r2o_trace_upconv.txt 2>&1
grep -E "==|kernel time|^[01] " gpurun_out/r2o_trace_upconv.txt
timeout 300 python scripts/profile_ops.py 5 > gpurun_out/r2o_profile_ops.txt 2>&1; grep -E "^0/.*(up|dec)|serial" gpurun_out/r2o_profile_ops.txt
timeout 600 python bench.py --steps 200 --no-cpu-baseline > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2o_bench.json')); print(d['value'], d['e2e']['value'], d['e2e']['p50_latency_ms'], d['config']['gflop_per_frame_executed'])
for s in d['roofline']['stages']:
    if 'conv' in s['kernel']: print(s['kernel'], s['launches_per_frame'], round(s['us_per_frame'],1), round(s['achieved'],1))"
