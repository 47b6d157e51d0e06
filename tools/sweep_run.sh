#!/bin/bash
# run tools/quick_bench.py against every tuning variant built by tools/sweep_build.py
cd "$(dirname "$0")/.."
for lib in tools/_sweep/lib_*.so; do
  echo "== $lib"
  TCL_B200_LIB=$PWD/$lib python tools/quick_bench.py 2>&1 | tail -5
done
