#!/bin/bash
# checks and times the built configurations of the blocked-sweep K2 kernel (matrices per warp and stage, warps per CTA, stages, unroll)
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_zsolve.py -q -k "sweep" 2>&1 | grep -E "^E  |passed|failed|Error" | head -20
python tools/bench_k2.py 32 1250000 gj sweep:4,9,1,4 sweep:4,9,1,2 sweep:4,9,1,1 sweep:2,15,1,2 sweep:2,15,1,1 sweep:2,9,2,2 sweep:4,5,2,4 sweep:2,12,1,2 sweep:4,8,1,4 2>&1 | grep -E "q=|rror"
python tools/bench_k2.py 16 1000000 gj sweep:4,16,2,4 sweep:4,16,2,1 sweep:2,16,2,2 2>&1 | grep -E "q=|rror"
# shared-memory hazards between the in-place phases (small cases)
timeout 200 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_zsolve.py -x -q -k "sweep and (32-3- or 16-1-)" 2>&1 | grep -E "RACECHECK|azard|passed|failed|ERROR SUMMARY|rror" | head -12
