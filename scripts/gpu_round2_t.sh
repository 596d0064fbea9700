#!/bin/bash
# pass t: final verification of HEAD — full GPU suite, smoke, default bench line; layer-by-layer graph (VPB_UPCONV=0) parity
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=300 --tb=short 2>&1 | tail -6 > gpurun_out/r2t_pytest.log; cat gpurun_out/r2t_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
VPB_UPCONV=0 timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_real_images_gpu.py -m gpu -q --timeout=300 --tb=short -k "not multitask_shares" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2t_bench.json')); r=d['roofline']; print(d['value'], d['e2e']['value'], d['e2e']['p50_latency_ms'], d['steps'], d['timed_region_s'], r['kernel'], r['achieved'], r['frac'], d['cpu_baseline']['value'], d['config']['frames_in_flight_per_gpu'])"
