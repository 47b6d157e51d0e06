#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_v20.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest_v20.log
tail -5 gpurun_out/r2_pytest_v20.log
python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/plain_bench_r02.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/ncu_launches_r02.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:fused_forward_ws -c 1 -f -o /tmp/prof_bench_r02 python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/ncu_full_r02.log 2>&1
python tools/ncu_summary.py /tmp/prof_bench_r02.ncu-rep gpurun_out/r02_bench_ws_kernel_sintel1041.txt > /dev/null 2>&1
python tools/ncu_traffic.py /tmp/prof_bench_r02.ncu-rep sintel_full 1041 > gpurun_out/ncu_traffic_r02.log 2>&1; cp profiles/r02_bench_traffic.json gpurun_out/ 2>/dev/null
python tools/ncu_lines.py /tmp/prof_bench_r02.ncu-rep 14572992 40 > gpurun_out/r02_bench_ws_kernel_lines.txt 2>&1
for k in upsample_flow cv2_fb_check cv2_remap hwc_split fused_forward_generic warp_backward; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o /tmp/prof_r02_$k python tools/bench_ops.py > gpurun_out/ncu_ops_$k.log 2>&1
  python tools/ncu_summary.py /tmp/prof_r02_$k.ncu-rep gpurun_out/r02_${k}.txt > /dev/null 2>&1
done
ls -la gpurun_out/ | tail -20; du -sh gpurun_out
