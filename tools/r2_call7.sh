#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_v13.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest_v13.log
tail -4 gpurun_out/r2_pytest_v13.log
timeout 300 python tools/bench_ops.py > gpurun_out/r2_bench_ops_v13.txt 2>&1; tail -30 gpurun_out/r2_bench_ops_v13.txt
timeout 300 python tools/quick_bench.py > gpurun_out/r2_quick_v13.txt 2>&1; cat gpurun_out/r2_quick_v13.txt
timeout 300 python tools/small_launch.py > gpurun_out/r2_small_v13.txt 2>&1; cat gpurun_out/r2_small_v13.txt
