#!/bin/bash
N=$1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tools/e2e_probe.py 500 2>&1 | grep "N=" 
