#!/bin/bash
# q = 64: the blocked-sweep K2 kernel against the blocked Cholesky kernel (K2 alone, engine tests, one sweep of the config-4 shape)
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_zsolve.py -q -k "sweep" 2>&1 | grep -E "^E  |passed|failed|Error" | head -20
python tools/bench_k2.py 64 400000 blocked sweep:1,6,1,1 sweep:1,4,2,1 sweep:2,4,1,2 sweep:2,4,1,1 2>&1 | grep -E "q=|rror"
PYVB_K2=sweep python -m pytest tests/test_gpu_i8.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -k "64 or ard or shape6 or shape7 or additive" 2>&1 | grep -E "^E  |passed|failed|Error" | head -20
python tools/sweep_time.py 400000 512 64 0.3 2>&1 | tail -1
PYVB_K2=sweep python tools/sweep_time.py 400000 512 64 0.3 2>&1 | tail -1
