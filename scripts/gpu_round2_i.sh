#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=300 --tb=short 2>&1 | tail -12 > gpurun_out/r2i_pytest.log; cat gpurun_out/r2i_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 200 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2i_bench.json')); print(d['value'], d['e2e']['value'], d['e2e']['p50_latency_ms'], d['roofline']['achieved'], d['roofline']['frac'])
for s in d['roofline']['stages']: print(s['kernel'], round(s['us_per_frame'],1), round(s['frac'],3))"
timeout 600 python bench.py --autospeed --steps 200 --no-cpu-baseline > gpurun_out/r2i_bench_autospeed.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2i_bench_autospeed.json')); print('autospeed', d['value'], d['e2e'])"
