#!/bin/bash
# pass v: engine construction after the upload-ordering fix — smoke, engine parity tests, one bench line
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_upconv_gpu.py -m gpu -q --timeout=300 --tb=short 2>&1 | tail -3
timeout 300 python bench.py --steps 200 --no-cpu-baseline > gpurun_out/r2v_bench.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2v_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['frac_executed'])"
