#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_v46.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest_v46.log
tail -4 gpurun_out/r2_pytest_v46.log
timeout 300 python tools/small_launch.py 2>&1 | grep -E "mask (direct|tma)" | tee gpurun_out/r2_small_v46.txt
