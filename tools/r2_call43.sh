#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_v43.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest_v43.log
tail -4 gpurun_out/r2_pytest_v43.log
timeout 300 python tools/debug_inf.py 2>&1 | tee gpurun_out/r2_debug_inf.txt
timeout 300 python tools/small_launch.py 2>&1 | tee gpurun_out/r2_small_v43.txt
