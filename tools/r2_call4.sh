#!/bin/bash
mkdir -p gpurun_out
export TCL_B200_LIB=$PWD/tools/_sweep/lib_hot_w16.so
python tools/prof_hot.py 256 5 > gpurun_out/plain_prof.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fused_forward_ws -s 3 -c 1 -f -o gpurun_out/prof_v11_w16 python tools/prof_hot.py 256 5 > gpurun_out/ncu_v11.log 2>&1
tail -3 gpurun_out/ncu_v11.log
