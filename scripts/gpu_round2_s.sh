#!/bin/bash
# pass s: measured 16-bit deviations (for the gates) + frames-in-flight sweep on the fused graph
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_engine_gpu.py -m gpu -q -s --timeout=200 -k "bf16 or single_model_parity" 2>&1 | grep -E "sigma|passed|failed" | head -20
for n in 2 4 5; do
  timeout 300 python bench.py --steps 200 --no-cpu-baseline --inflight $n > gpurun_out/r2s_bench_inflight$n.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/r2s_bench_inflight$n.json')); print('inflight $n', d['value'], d['e2e']['value'], d['e2e']['p50_latency_ms'])"
done
