#!/bin/bash
# pass p: full verification of the fused-upconv build + refreshed profiles
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=300 --tb=short 2>&1 | tail -12 > gpurun_out/r2p_pytest.log; cat gpurun_out/r2p_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 200 > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2p_bench.json')); r=d['roofline']; print(d['value'], d['e2e']['value'], d['e2e']['p50_latency_ms'], r['kernel'], r['achieved'], r['frac'], r['achieved_reference_equivalent'], d['cpu_baseline'])
for s in r['stages']: print(s['kernel'], s['launches_per_frame'], round(s['us_per_frame'],1), round(s['achieved'],1), round(s['frac'],3))"
timeout 600 python bench.py --autospeed --steps 200 --no-cpu-baseline > gpurun_out/r2p_bench_autospeed.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2p_bench_autospeed.json')); print('autospeed', d['value'], d['e2e'])"
bash scripts/ncu_launch_list.sh > gpurun_out/r2p_launch_list.txt 2>&1; head -30 gpurun_out/r2p_launch_list.txt
bash scripts/ncu_conv_traffic.sh upconv_pair 60 60 > gpurun_out/r2p_upconv_traffic.txt 2>&1; tail -2 gpurun_out/r2p_upconv_traffic.txt; cp gpurun_out/conv_traffic.json gpurun_out/r2p_upconv_traffic.json
exp() { ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/$1.csv 2>/dev/null; }
timeout 300 ncu --set full --clock-control none -k regex:'upconv_pair' -s 1 -c 1 -f -o /tmp/r2_ncu_upconv_up3dec6 python scripts/trace_upconv.py 80 160 256 256 32 > gpurun_out/r2_ncu_upconv_up3dec6.log 2>&1; exp r2_ncu_upconv_up3dec6
timeout 300 ncu --set full --clock-control none -k regex:'upconv_pair' -s 1 -c 1 -f -o /tmp/r2_ncu_upconv_up4dec8 python scripts/trace_upconv.py 160 320 128 128 0 > gpurun_out/r2_ncu_upconv_up4dec8.log 2>&1; exp r2_ncu_upconv_up4dec8
ls -la gpurun_out/*.csv | tail -5; du -sh gpurun_out
