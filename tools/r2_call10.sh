#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_v17.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest_v17.log
tail -3 gpurun_out/r2_pytest_v17.log
: > gpurun_out/r2_sustained_v17.txt
for v in r01hot hot_w16 hot_w8; do
  TCL_B200_LIB=$PWD/tools/_sweep/lib_$v.so timeout 120 python tools/sustained.py 3 >> gpurun_out/r2_sustained_v17.txt 2>&1
done
cat gpurun_out/r2_sustained_v17.txt | tail -4
timeout 300 python tools/bench_ops.py > gpurun_out/r2_bench_ops_v17.txt 2>&1; grep "fused, ff\|(mask)\|bf16 frames" gpurun_out/r2_bench_ops_v17.txt
timeout 200 python tools/quick_bench.py > gpurun_out/r2_quick_v17.txt 2>&1; cat gpurun_out/r2_quick_v17.txt
