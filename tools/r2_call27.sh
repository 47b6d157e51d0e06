#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2_sustained_v27.txt
for v in hot_bh44 hot_nb3_bh44 hot_nb3_bh48 hot_nb3_bh49_s3 hot_bh44; do
  TCL_B200_LIB=$PWD/tools/_sweep/lib_$v.so timeout 120 python tools/sustained.py 3 >> gpurun_out/r2_sustained_v27.txt 2>&1
done
cat gpurun_out/r2_sustained_v27.txt | tail -9
