#!/bin/bash
# checks and times the built configurations of the Gauss-Jordan K2 kernel (rows per lane, warps per CTA)
cd "$(dirname "$0")/.."
for cfg in 4,8 2,16; do
  echo "== q=16 PYVB_GJ=$cfg"
  PYVB_GJ=$cfg python -m pytest tests/test_gpu_zsolve.py -x -q -k "gj and 16" 2>&1 | tail -1
  PYVB_GJ=$cfg python tools/bench_k2.py 16 1000000 gj 2>&1 | tail -1
done
for cfg in 2,8 1,12; do
  echo "== q=32 PYVB_GJ=$cfg"
  PYVB_GJ=$cfg python -m pytest tests/test_gpu_zsolve.py -x -q -k "gj and 32" 2>&1 | tail -1
  PYVB_GJ=$cfg python tools/bench_k2.py 32 1250000 gj 2>&1 | tail -1
done
