#!/bin/bash
# pass u: the default bench line with roofline.achieved in algorithmic FLOPs and the executed figure beside it
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2u_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2u_bench.json')); r=d['roofline']; print(d['value'], d['e2e']['value'], d['e2e']['p50_latency_ms'], r['kernel'], r['achieved'], r['frac'], r['achieved_executed'], r['frac_executed'], r['frac_vs_burst_peak'])
for s in r['stages']:
    if s['bound']=='tensor': print(s['kernel'], round(s['us_per_frame'],1), round(s['achieved'],1), round(s['frac'],3), s.get('achieved_executed'))"
