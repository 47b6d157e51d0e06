#!/bin/bash
# round-2 evidence for profiles/: tests, launch list, full ncu capture of the dominant kernel (summary + traffic JSON, written before
# the bench run so that its line carries roofline.traffic for exactly these kernel sources), the bench line, the other kernels, op tables
mkdir -p gpurun_out/p
P=gpurun_out/p
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $P/r02_gpu.txt
timeout 400 python -m pytest tests -m gpu -x -q > $P/r02_pytest_gpu.log 2>&1; echo "pytest rc $?" >> $P/r02_pytest_gpu.log; tail -3 $P/r02_pytest_gpu.log
python bench.py --steps 3 --warmup 3 --no-extras > $P/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $P/r02_bench_launches.csv python bench.py --steps 3 --warmup 3 --no-extras > $P/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:fused_forward_ws -c 1 -f -o /tmp/prof_bench_r02 python bench.py --steps 3 --warmup 3 --no-extras > $P/ncu_full.log 2>&1
python tools/ncu_summary.py /tmp/prof_bench_r02.ncu-rep $P/r02_bench_ws_kernel_sintel1041.txt > /dev/null 2>&1
python tools/ncu_traffic.py /tmp/prof_bench_r02.ncu-rep sintel_full 1041 > $P/ncu_traffic.log 2>&1; cp profiles/r02_bench_traffic.json $P/ 2>/dev/null
python tools/ncu_lines.py /tmp/prof_bench_r02.ncu-rep 14572992 40 > $P/r02_bench_ws_kernel_lines.txt 2>&1
timeout 900 python bench.py > $P/r02_bench_line.json 2> $P/r02_bench_line.err; echo "bench rc $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $P/r02_bench_line_reference_arm.json 2> $P/r02_bench_ref.err; echo "reference arm rc $?"
python tools/small_launch.py one > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fused_forward_direct -s 1 -c 1 -f -o /tmp/prof_r02_direct python tools/small_launch.py one > $P/ncu_direct.log 2>&1
python tools/ncu_summary.py /tmp/prof_r02_direct.ncu-rep $P/r02_fused_forward_direct_b16.txt > /dev/null 2>&1
for k in upsample_flow cv2_fb_check cv2_remap hwc_split fused_forward_generic warp_backward reconet_forward ruder_input; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o /tmp/prof_r02_$k python tools/prof_ops.py > $P/ncu_ops_$k.log 2>&1
  python tools/ncu_summary.py /tmp/prof_r02_$k.ncu-rep $P/r02_${k}.txt > /dev/null 2>&1
done
timeout 300 python tools/bench_ops.py > $P/r02_bench_ops.txt 2>&1
timeout 200 python tools/window_bench.py > $P/r02_window_bench.txt 2>&1
timeout 200 python tools/small_launch.py > $P/r02_small_launch.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $P/r02_smoke.log 2>&1; tail -1 $P/r02_smoke.log
rm -f $P/ncu_ops_*.log $P/plain_bench.log
ls -la $P | tail -32; du -sh gpurun_out
