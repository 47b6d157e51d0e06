#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/small_launch.py variants 2>&1 | tee gpurun_out/r2_small_v42.txt
for v in bf16_bh42 bf16_bh48 bf16_bh54 bf16_bh60; do
  TCL_B200_LIB=$PWD/tools/_sweep/lib_$v.so timeout 120 python tools/quick_bf16.py 2>&1 | tail -2 | tee -a gpurun_out/r2_bf16_v42.txt
done
