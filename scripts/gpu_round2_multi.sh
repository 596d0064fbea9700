#!/bin/bash
# 8-GPU pass: config 5 (C++ ncclAllGather + Estimator fusion) test at world = 8, config-5 bench at N = 2, 4, 8
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m pytest tests/test_multicam_gpu.py -m gpu -q -s --timeout=500 --tb=short 2>&1 | tail -15 > gpurun_out/r2m_multicam_test.log; cat gpurun_out/r2m_multicam_test.log
for n in 2 4 8; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520 + n)) \
     bench.py --gpus $n --steps 100 --config5 > gpurun_out/r2m_config5_${n}gpu.json 2> gpurun_out/r2m_config5_${n}gpu.err
  echo "config5 N=$n rc=$?"; python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/r2m_config5_${n}gpu.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['allgather'])
except Exception as e: print('parse error', e)"
done
