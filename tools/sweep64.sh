#!/bin/bash
# the blocked-sweep K2 kernel (PYVB_K2=sweep): K2 alone against LAPACK, the engine tests that run q = 64 / ARD, timing against the
# Gauss-Jordan (q = 32) and blocked Cholesky (q = 64) kernels, one sweep of the config-4 shape
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_zsolve.py -q -k "sweep" 2>&1 | grep -E "^E  |passed|failed|Error" | head -20
PYVB_K2=sweep python -m pytest tests/test_gpu_i8.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -k "64 or ard or shape6 or shape7 or additive or 32" 2>&1 | grep -E "^E  |passed|failed|Error" | head -20
python tools/bench_k2.py 64 400000 blocked sweep:1,6,1,1 sweep:2,4,1,1 2>&1 | grep -E "q=|rror"
python tools/bench_k2.py 32 1250000 gj sweep:4,8,1,4 sweep:4,9,1,1 2>&1 | grep -E "q=|rror"
python tools/sweep_time.py 400000 512 64 0.3 2>&1 | tail -n 1
PYVB_K2=sweep python tools/sweep_time.py 400000 512 64 0.3 2>&1 | tail -n 1
