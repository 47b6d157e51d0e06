#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 300 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q > gpurun_out/r2_pytest_multirank.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest_multirank.log
tail -4 gpurun_out/r2_pytest_multirank.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r02_n2.json 2> gpurun_out/bench_r02_n2.err; echo "bench rc $?"; tail -5 gpurun_out/bench_r02_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_n2.json').read().strip().splitlines()[-1])
for k in ('value','gpix_per_s','ms_per_step','n_gpus','gpu_launches','clocks','result_check'): print(k, d.get(k))
print('roofline', {k:d['roofline'][k] for k in ('achieved','frac','kernel_ms_per_launch')})
print('e2e', {k:v for k,v in d['e2e'].items() if 'note' not in k})
print('strong', json.dumps(d.get('strong_scaling'), indent=1))
PY
