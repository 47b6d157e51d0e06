#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_v23.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest_v23.log
tail -4 gpurun_out/r2_pytest_v23.log
timeout 200 python tools/sustained.py 3
timeout 200 python tools/quick_bench.py
timeout 300 python tools/bench_ops.py > gpurun_out/r2_bench_ops_v23.txt 2>&1; grep "cv2compat\|fused, ff\|(mask)\|bf16 frames" gpurun_out/r2_bench_ops_v23.txt | head -12
