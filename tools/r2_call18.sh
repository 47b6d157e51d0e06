#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_v22.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest_v22.log
tail -6 gpurun_out/r2_pytest_v22.log
timeout 200 python tools/window_bench.py > gpurun_out/r2_window_v22.txt 2>&1; cat gpurun_out/r2_window_v22.txt
timeout 300 python tools/bench_ops.py > gpurun_out/r2_bench_ops_v22.txt 2>&1; grep "cv2compat\|fused, ff\|(mask)\|bf16 frames\|fwd+bwd" gpurun_out/r2_bench_ops_v22.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_c.json 2> gpurun_out/bench_r02_c.err; echo "bench rc $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_c.json').read().strip().splitlines()[-1])
for k in ('value','gpix_per_s','ms_per_step','clocks'): print(k, d.get(k))
print('roofline', {k:d['roofline'][k] for k in ('achieved','frac','kernel_ms_per_launch','traffic')}, d['roofline']['burst']['frac'])
for w in d.get('other_workloads',[]): print(w.get('workload'), w.get('gpix_per_s'), w.get('frac_of_measured_peak'), w.get('ms_per_launch_median'), w.get('ms_per_step_device'), w.get('ms_per_step_eager_autograd'), w.get('error'))
PY
