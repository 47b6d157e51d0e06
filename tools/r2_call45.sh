#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_v45.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest_v45.log
tail -4 gpurun_out/r2_pytest_v45.log
timeout 300 python tools/bench_ops.py > gpurun_out/r2_bench_ops_v45.txt 2>&1; head -12 gpurun_out/r2_bench_ops_v45.txt
