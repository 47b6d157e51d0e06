#!/bin/bash
# sustained throughput of every tuning variant built by tools/sweep_build.py
cd "$(dirname "$0")/.."
for lib in tools/_sweep/lib_hot*.so; do
  TCL_B200_LIB=$PWD/$lib python tools/sustained.py ${1:-3} 2>&1 | tail -1
  sleep 2
done
