#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/plain_bench_r02.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/ncu_launches_r02.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:fused_forward_ws -c 1 -f -o gpurun_out/prof_bench_r02 python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/ncu_full_r02.log 2>&1
tail -2 gpurun_out/ncu_full_r02.log
for k in upsample_flow cv2_fb_check cv2_remap hwc_split fused_forward_generic warp_backward; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/prof_r02_$k python tools/bench_ops.py > gpurun_out/ncu_ops_$k.log 2>&1
  tail -1 gpurun_out/ncu_ops_$k.log
done
ls -la gpurun_out/*.ncu-rep | tail -8
