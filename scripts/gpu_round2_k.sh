#!/bin/bash
mkdir -p gpurun_out
python scripts/trace_ws.py 80 160 256 256 > gpurun_out/r2k_trace_ws_up3.txt 2>&1; cat gpurun_out/r2k_trace_ws_up3.txt
python scripts/trace_ws.py 160 320 128 128 > gpurun_out/r2k_trace_ws_up4.txt 2>&1; cat gpurun_out/r2k_trace_ws_up4.txt
