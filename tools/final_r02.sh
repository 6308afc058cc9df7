#!/bin/bash
# final verification of the round-2 tree on one B200: GPU tests, the bench line, smoke, the launch list of the bench command and
# one full ncu capture of the dominant kernel (K2, Gauss-Jordan) at the two bench sizes
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r02_pytest_gpu_1gpu.txt; tail -1 gpurun_out/r02_pytest_gpu_1gpu.txt
python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench rc=$?"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py --steps 3 --warmup 3 --quick > gpurun_out/r02_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --quick > gpurun_out/r02_ncu_launch.log 2>&1
bash tools/ncu_k2.sh 16 1000000 gj r02_k2_gj16
bash tools/ncu_k2.sh 32 1250000 gj r02_k2_gj32
tail -1 gpurun_out/r02_k2_gj16_plain.log gpurun_out/r02_k2_gj32_plain.log
