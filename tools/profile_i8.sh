#!/bin/bash
# ncu evidence for the INT8-path sweep (1 GPU): launch list of the bench command + full captures of its six hot kernels
cd "$(dirname "$0")/.."
set -x
python bench.py --steps 2 --warmup 3 --no-cpu --no-f32 > gpurun_out/plain_bench_i8.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01_i8.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-f32 > gpurun_out/ncu_launch_i8.log 2>&1
K='regex:zstep_i8|zstep_dmma|zsolve|digitize|stats_i8|stats_dmma'
python tools/time_i8.py --only i8 151552 256 16 0.2 > gpurun_out/plain_i8_c2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "$K" -s 12 -c 6 -f -o gpurun_out/prof_r01_i8_c2 \
    python tools/time_i8.py --only i8 151552 256 16 0.2 > gpurun_out/ncu_i8_c2.log 2>&1
python tools/time_i8.py --only i8 151552 1024 32 0.3 > gpurun_out/plain_i8_c3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "$K" -s 12 -c 6 -f -o gpurun_out/prof_r01_i8_c3 \
    python tools/time_i8.py --only i8 151552 1024 32 0.3 > gpurun_out/ncu_i8_c3.log 2>&1
ls -la gpurun_out/prof_r01_i8_c*.ncu-rep
