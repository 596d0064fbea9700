#!/bin/bash
# pass q (2 GPUs): new upconv cases, two-device / multicam tests, 2-GPU weak scaling of the main bench, config 5 at N = 2
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_upconv_gpu.py -m gpu -q --timeout=120 --tb=short 2>&1 | tail -4
timeout 600 python -m pytest tests/test_multicam_gpu.py -m gpu -q --timeout=400 --tb=short 2>&1 | tail -4
for mode in main config5; do
  extra=""; [ $mode = config5 ] && extra="--config5"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
     bench.py --gpus 2 --steps 100 --no-cpu-baseline $extra > gpurun_out/r2q_${mode}_2gpu.json 2> gpurun_out/r2q_${mode}_2gpu.err
  echo "$mode N=2 rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/r2q_${mode}_2gpu.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d.get('e2e',{}).get('value'), d.get('allgather'))"
done
