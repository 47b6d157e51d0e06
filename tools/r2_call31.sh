#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_v28.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest_v28.log
tail -4 gpurun_out/r2_pytest_v28.log
timeout 200 python tools/sustained.py 3
timeout 200 python tools/quick_bench.py
timeout 300 python tools/bench_ops.py > gpurun_out/r2_bench_ops_v28.txt 2>&1; sed -n 10,36p gpurun_out/r2_bench_ops_v28.txt
