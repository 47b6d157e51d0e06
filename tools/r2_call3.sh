#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2_sustained_diag.txt
for v in hot_w16 diag1_w16 diag2_w16 w8_th16_ns2 w8_th16_ns3 w8_th16_ns4; do
  TCL_B200_LIB=$PWD/tools/_sweep/lib_$v.so timeout 120 python tools/sustained.py 2 >> gpurun_out/r2_sustained_diag.txt 2>&1
done
cat gpurun_out/r2_sustained_diag.txt | tail -12
