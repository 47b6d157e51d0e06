#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_v14.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest_v14.log
tail -3 gpurun_out/r2_pytest_v14.log
timeout 300 python tools/bench_ops.py > gpurun_out/r2_bench_ops_v14.txt 2>&1; sed -n 10,60p gpurun_out/r2_bench_ops_v14.txt
echo "---- packed_given=0"
TCL_B200_LIB=$PWD/tools/_sweep/lib_full_pg0.so timeout 300 python tools/bench_ops.py > gpurun_out/r2_bench_ops_v14_pg0.txt 2>&1; grep "mask)" gpurun_out/r2_bench_ops_v14_pg0.txt
timeout 300 python tools/small_launch.py > gpurun_out/r2_small_v14.txt 2>&1; cat gpurun_out/r2_small_v14.txt
