#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_v18.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest_v18.log
tail -3 gpurun_out/r2_pytest_v18.log
timeout 300 python tools/bench_ops.py > gpurun_out/r2_bench_ops_v18.txt 2>&1; grep "fused, ff\|(mask)\|bf16\|blend\|outputs\|fwd+bwd" gpurun_out/r2_bench_ops_v18.txt
echo "---- CW=16 everywhere"
TCL_B200_LIB=$PWD/tools/_sweep/lib_full_w16.so timeout 300 python tools/bench_ops.py > gpurun_out/r2_bench_ops_v18_w16.txt 2>&1; grep "fused, ff\|(mask)\|bf16 frames" gpurun_out/r2_bench_ops_v18_w16.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_a.json 2> gpurun_out/bench_r02_a.err; echo "bench rc $?"; tail -3 gpurun_out/bench_r02_a.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_a.json').read().strip().splitlines()[-1])
for k in ('value','gpix_per_s','ms_per_step','preheat_steps','gpu_launches','clocks'): print(k, d.get(k))
print('roofline', {k:d['roofline'][k] for k in ('achieved','frac','kernel_ms_per_launch','traffic')}, d['roofline']['burst'])
print('e2e', {k:v for k,v in d['e2e'].items() if k!='note'})
print('eager', d.get('cuda_eager_baseline'))
print('cpu', d.get('cpu_baseline'))
for w in d.get('other_workloads',[]): print(w.get('workload'), w.get('gpix_per_s'), w.get('frac_of_measured_peak'), w.get('ms_per_launch_median'), w.get('ms_per_step_device'), w.get('ms_per_step_eager_autograd'), w.get('error'))
PY
