#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=300 --durations=8 --tb=short 2>&1 | tail -40 > gpurun_out/r2h_pytest.log; cat gpurun_out/r2h_pytest.log
timeout 600 python scripts/diag_taps.py scene_seg 2>&1 | grep -E "^fp16|^fp32" > gpurun_out/r2h_diag_scene_seg.txt; cat gpurun_out/r2h_diag_scene_seg.txt
timeout 600 python bench.py --steps 200 --no-cpu-baseline > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2h_bench.json')); print(d['value'], d['e2e']['value'], d['e2e']['p50_latency_ms'])
for s in d['roofline']['stages']: print(s['kernel'], round(s['us_per_frame'],1))"
