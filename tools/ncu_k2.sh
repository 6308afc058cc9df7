#!/bin/bash
# one full ncu capture of a K2 kernel: tools/ncu_k2.sh <q> <N> <impl> <tag>   (report + raw csv into gpurun_out/)
cd "$(dirname "$0")/.."
q=$1; N=$2; impl=$3; tag=$4
python tools/bench_k2.py $q $N $impl > gpurun_out/${tag}_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:zsolve -s 1 -c 1 -f -o gpurun_out/$tag \
    python tools/bench_k2.py $q $N $impl > gpurun_out/${tag}_ncu.log 2>&1
ls -la gpurun_out/$tag.ncu-rep
