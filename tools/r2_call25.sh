#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_parity.py -k "host_entry" -m gpu -x -q 2>&1 | tail -2
timeout 300 python tools/e2e_probe.py 400 2>&1 | grep "N="
