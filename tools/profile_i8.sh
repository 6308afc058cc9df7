#!/bin/bash
# ncu evidence for the INT8-path sweep (1 GPU): launch list of the bench command + one full capture of every kernel of a sweep
cd "$(dirname "$0")/.."
set -x
python bench.py --steps 2 --warmup 3 --no-cpu --no-f32 > gpurun_out/plain_bench_i8.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01_i8.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-f32 > gpurun_out/ncu_launch_i8.log 2>&1
python tools/profile_sweep.py 151552 256 16 0.2 > gpurun_out/plain_i8_c2.log 2>&1 &&
ncu --profile-from-start off --set full --clock-control none --import-source on -f -o gpurun_out/prof_r01_i8_c2 \
    python tools/profile_sweep.py 151552 256 16 0.2 > gpurun_out/ncu_i8_c2.log 2>&1
python tools/profile_sweep.py 151552 1024 32 0.3 > gpurun_out/plain_i8_c3.log 2>&1 &&
ncu --profile-from-start off --set full --clock-control none --import-source on -f -o gpurun_out/prof_r01_i8_c3 \
    python tools/profile_sweep.py 151552 1024 32 0.3 > gpurun_out/ncu_i8_c3.log 2>&1
ls -la gpurun_out/prof_r01_i8_c*.ncu-rep
