#!/bin/bash
# first GPU pass of round 2 (2 GPUs): full parity suite incl. config 5 over NCCL, bench, config-5 bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 200 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --config5 > gpurun_out/r2a_config5_2gpu.json 2> gpurun_out/r2a_config5_2gpu.err; echo "config5 rc=$?"
tail -c 1500 gpurun_out/r2a_config5_2gpu.json; tail -3 gpurun_out/r2a_config5_2gpu.err
