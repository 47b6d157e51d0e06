#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_v41.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest_v41.log
tail -4 gpurun_out/r2_pytest_v41.log
timeout 300 python tools/small_launch.py variants 2>&1 | tee gpurun_out/r2_small_v41.txt
python tools/small_launch.py one > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fused_forward_direct -s 1 -c 1 -f -o /tmp/prof_direct python tools/small_launch.py one > gpurun_out/ncu_direct.log 2>&1
python tools/ncu_summary.py /tmp/prof_direct.ncu-rep gpurun_out/r2_direct_b16.txt > /dev/null 2>&1
head -45 gpurun_out/r2_direct_b16.txt
