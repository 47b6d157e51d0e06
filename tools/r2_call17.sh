#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_v21.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest_v21.log
tail -6 gpurun_out/r2_pytest_v21.log
timeout 200 python tools/quick_bench.py > gpurun_out/r2_quick_v21.txt 2>&1; cat gpurun_out/r2_quick_v21.txt
timeout 200 python tools/window_bench.py > gpurun_out/r2_window_v21.txt 2>&1; cat gpurun_out/r2_window_v21.txt
