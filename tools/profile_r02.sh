#!/bin/bash
# ncu evidence of round 2 (1 GPU), at the BENCH sizes: launch list of the bench command + one full capture of every kernel of a
# steady-state sweep at config 2 (N = 1M, D = 256, q = 16) and at the config-3 shard (N = 1.25M, D = 1024, q = 32).
# The .ncu-rep files (35 / 50 MB) stay on the box: their raw pages are exported to CSV there (gpurun_out/ is capped at 64 MiB).
cd "$(dirname "$0")/.."
set -x
python bench.py --steps 3 --warmup 3 --quick > gpurun_out/r02_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --quick > gpurun_out/r02_ncu_launch.log 2>&1
for cfg in "c2 1000000 256 16 0.2" "c3 1250000 1024 32 0.3"; do
    set -- $cfg
    python tools/profile_sweep.py $2 $3 $4 $5 > gpurun_out/r02_plain_$1.log 2>&1 &&
    ncu --profile-from-start off --set full --clock-control none -f -o /tmp/r02_prof_$1 \
        python tools/profile_sweep.py $2 $3 $4 $5 > gpurun_out/r02_ncu_$1.log 2>&1
    ncu -i /tmp/r02_prof_$1.ncu-rep --page raw --csv > gpurun_out/r02_prof_$1_raw.csv 2>/dev/null
done
ls -la gpurun_out/r02_prof_*_raw.csv gpurun_out/r02_launches_bench.csv
