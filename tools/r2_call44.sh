#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_v44.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest_v44.log
tail -4 gpurun_out/r2_pytest_v44.log
timeout 300 python tools/debug_inf.py 2>&1 | grep -c inf
for lib in hot_bh44 hot hot_bh44 hot; do
  TCL_B200_LIB=$PWD/tools/_sweep/lib_$lib.so timeout 100 python tools/sustained.py 3 2>&1 | tail -1 | tee -a gpurun_out/r2_sustained_v44.txt
  sleep 2
done
