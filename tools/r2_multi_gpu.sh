#!/bin/bash
# usage: r2_multi_gpu.sh N
N=$1
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
timeout 300 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_parity.py -k "multirank or two_rank or band_mode" -m gpu -x -q > gpurun_out/r2_pytest_n$N.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest_n$N.log
tail -3 gpurun_out/r2_pytest_n$N.log
fi
for wl in sintel_full hd1080_window uhd4k_stress; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 --workload $wl > gpurun_out/bench_r02_${wl}_n$N.json 2> gpurun_out/bench_r02_${wl}_n$N.err; echo "$wl bench rc $?"
  python - $wl $N <<'PY'
import json, sys
wl, n = sys.argv[1], sys.argv[2]
try:
    d=json.loads(open(f'gpurun_out/bench_r02_{wl}_n{n}.json').read().strip().splitlines()[-1])
except Exception as ex:
    print("no line", ex); print(open(f'gpurun_out/bench_r02_{wl}_n{n}.err').read()[-1500:]); sys.exit(0)
print(wl, 'value', d['value'], 'gpix', d['gpix_per_s'], 'ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'], 'clk', d['clocks']['sm_mhz'])
e=d.get('e2e',{}); print('  e2e', {k:e.get(k) for k in ('value','h2d_gb_per_s','h2d_ceiling_gb_per_s','h2d_frac_of_ceiling','error')})
if 'strong_scaling' in d: s=d['strong_scaling']; print('  strong', {k:s.get(k) for k in ('value','ms_per_step','share_outside_fused_kernel','error')}, s.get('result_check'))
if 'band_split' in d: print('  band', d['band_split'])
PY
done
