#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/diag_taps.py scene_seg > gpurun_out/r2g_diag_scene_seg.txt 2>&1; cat gpurun_out/r2g_diag_scene_seg.txt | grep -v Warn | tail -12
timeout 600 python bench.py --autospeed --steps 200 > gpurun_out/r2g_bench_autospeed.json 2> gpurun_out/r2g_bench_autospeed.err; echo "bench autospeed rc=$?"; tail -c 1200 gpurun_out/r2g_bench_autospeed.json; tail -3 gpurun_out/r2g_bench_autospeed.err
