#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2_sustained_v13.txt
for v in r01hot hot_w16_scalar hot_w16 hot hot_scalar r01hot hot_w16; do
  TCL_B200_LIB=$PWD/tools/_sweep/lib_$v.so timeout 120 python tools/sustained.py 3 >> gpurun_out/r2_sustained_v13.txt 2>&1
done
cat gpurun_out/r2_sustained_v13.txt | tail -8
