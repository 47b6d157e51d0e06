#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_v11.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest_v11.log
tail -15 gpurun_out/r2_pytest_v11.log
: > gpurun_out/r2_sustained_v11.txt
timeout 120 python tools/sustained.py 3 >> gpurun_out/r2_sustained_v11.txt 2>&1
for v in hot hot_w16 hot_bw76 hot_w16_bw76; do
  TCL_B200_LIB=$PWD/tools/_sweep/lib_$v.so timeout 120 python tools/sustained.py 3 >> gpurun_out/r2_sustained_v11.txt 2>&1
done
cat gpurun_out/r2_sustained_v11.txt | tail -12
TCL_B200_LIB=$PWD/tools/_sweep/lib_trace.so timeout 120 python tools/trace_pipeline.py 64 > gpurun_out/r2_trace_v11.txt 2>&1
tail -3 gpurun_out/r2_trace_v11.txt
