#!/bin/bash
# pass r: final bench line (with the committed upconv traffic figure) + launch list of the fused graph
mkdir -p gpurun_out
timeout 900 python bench.py --steps 200 > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2r_bench.json')); r=d['roofline']; print(d['value'], d['e2e']['value'], d['e2e']['p50_latency_ms'], r['kernel'], r['achieved'], r['frac'], r['traffic'], r['traffic_src'])"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2r_bench_reference.json 2>/dev/null; tail -c 600 gpurun_out/r2r_bench_reference.json
bash scripts/ncu_launch_list.sh > gpurun_out/r2r_launch_list.txt 2>&1; head -34 gpurun_out/r2r_launch_list.txt
