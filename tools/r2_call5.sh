#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2_sustained_v12.txt
for v in hot hot_w16; do
  TCL_B200_LIB=$PWD/tools/_sweep/lib_$v.so timeout 120 python tools/sustained.py 3 >> gpurun_out/r2_sustained_v12.txt 2>&1
done
cat gpurun_out/r2_sustained_v12.txt | tail -4
export TCL_B200_LIB=$PWD/tools/_sweep/lib_hot_w16.so
python tools/prof_hot.py 256 5 > gpurun_out/plain_prof.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fused_forward_ws -s 3 -c 1 -f -o gpurun_out/prof_v12_w16 python tools/prof_hot.py 256 5 > gpurun_out/ncu_v12.log 2>&1
tail -2 gpurun_out/ncu_v12.log
