#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gemm_gpu.py tests/test_engine_gpu.py -m gpu -q -x --timeout=300 --tb=short 2>&1 | tail -8
python scripts/trace_ws.py 80 160 256 256 > gpurun_out/r2l_trace_ws_up3.txt 2>&1; cat gpurun_out/r2l_trace_ws_up3.txt
timeout 300 python scripts/bench_conv.py lin > gpurun_out/r2l_bench_conv.txt 2>&1; grep -E "up[0-9]|total" gpurun_out/r2l_bench_conv.txt
timeout 600 python bench.py --steps 200 --no-cpu-baseline > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2l_bench.json')); print(d['value'], d['e2e']['value'], d['e2e']['p50_latency_ms'])
for s in d['roofline']['stages']:
    if 'conv' in s['kernel']: print(s['kernel'], s['launches_per_frame'], round(s['us_per_frame'],1), round(s['achieved'],1))"
