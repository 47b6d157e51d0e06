#!/bin/bash
# round-2 first GPU call: baseline + probes
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r2_gpu.txt
( cd tools/tma_probe && bash run.sh ) > gpurun_out/r2_tma_probe.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_start.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest_start.log
python tools/sustained.py 3 > gpurun_out/r2_sustained.txt 2>&1
TCL_B200_LIB=$PWD/tools/_sweep/lib_hot.so python tools/sustained.py 3 >> gpurun_out/r2_sustained.txt 2>&1
TCL_B200_LIB=$PWD/tools/_sweep/lib_hot_w8.so python tools/sustained.py 3 >> gpurun_out/r2_sustained.txt 2>&1
TCL_B200_LIB=$PWD/tools/_sweep/lib_trace.so python tools/trace_pipeline.py 64 > gpurun_out/r2_trace.txt 2>&1
tail -3 gpurun_out/r2_pytest_start.log; cat gpurun_out/r2_sustained.txt; cat gpurun_out/r2_tma_probe.txt; tail -5 gpurun_out/r2_trace.txt
