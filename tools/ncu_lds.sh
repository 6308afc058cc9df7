#!/bin/bash
# one full ncu capture of the LDS smoother kernel at config 5 (report into gpurun_out/): tools/ncu_lds.sh <tag>
cd "$(dirname "$0")/.."
tag=$1
python tools/bench_lds.py --no-cpu > gpurun_out/${tag}_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:lds_iterate -s 3 -c 1 -f -o gpurun_out/$tag \
    python tools/bench_lds.py --no-cpu > gpurun_out/${tag}_ncu.log 2>&1
ls -la gpurun_out/$tag.ncu-rep
