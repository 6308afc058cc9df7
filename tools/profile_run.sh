#!/bin/bash
# ncu evidence for the round: launch lists of whole sweeps + full captures of every hot kernel (1 GPU)
cd "$(dirname "$0")/.."
set -x
K64='regex:zstep_dmma|zsolve_tpm|zsolve_blocked|stats_dmma|stats_reduce'
K32='regex:zstep_f32|zsolve_tpm|stats_f32'
python bench.py --steps 2 --warmup 3 --no-cpu --no-f32 > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01b.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-f32 > gpurun_out/ncu_launch.log 2>&1
python tools/profile_kernels.py > gpurun_out/plain64.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "$K64" -s 4 -c 4 -f -o gpurun_out/prof_r01_f64_c2 \
    python tools/profile_kernels.py > gpurun_out/ncu64.log 2>&1
python tools/profile_kernels.py --precision f32 > gpurun_out/plain32.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01_f32.csv \
    python tools/profile_kernels.py --precision f32 > gpurun_out/ncu_launch32.log 2>&1
ncu --set full --clock-control none --import-source on -k "$K32" -s 3 -c 3 -f -o gpurun_out/prof_r01_f32_c2 \
    python tools/profile_kernels.py --precision f32 > gpurun_out/ncu32.log 2>&1
python tools/profile_kernels.py --N 151552 --D 1024 --q 32 > gpurun_out/plain64_c3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "$K64" -s 4 -c 4 -f -o gpurun_out/prof_r01_f64_c3 \
    python tools/profile_kernels.py --N 151552 --D 1024 --q 32 > gpurun_out/ncu64_c3.log 2>&1
ls -la gpurun_out/*.ncu-rep
