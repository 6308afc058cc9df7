#!/bin/bash
# the tile-swizzled blocked-sweep K2 kernel (PYVB_SWEEP_TILED=1): the whole GPU suite with it as the K2 of every q it covers, K2 alone
# against the defaults, one sweep of the config-4 shape and of the config-3 shard
cd "$(dirname "$0")/.."
PYVB_SWEEP_TILED=1 PYVB_K2=sweep python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  |passed|failed|Error" | head -12
export PYVB_SWEEP_TILED=1
python tools/bench_k2.py 64 400000 blocked sweep:2,4,1,1 sweep:1,6,1,1 2>&1 | grep -E "q=|rror"
python tools/bench_k2.py 32 1250000 gj sweep:4,8,1,1 sweep:4,8,1,4 sweep:2,12,1,1 sweep:2,8,2,1 2>&1 | grep -E "q=|rror"
python tools/bench_k2.py 16 1000000 gj sweep:4,12,2,4 2>&1 | grep -E "q=|rror"
python tools/sweep_time.py 400000 512 64 0.3 2>&1 | tail -n 1
PYVB_K2=sweep python tools/sweep_time.py 1250000 1024 32 0.3 2>&1 | tail -n 1
