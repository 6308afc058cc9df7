#!/bin/bash
# multi-GPU scaling runs on one box (fused peer-memory exchange): C2 per GPU (the bench default, weak scaling) at
# 2/4/8 GPUs, C3 (N=10M over 8 GPUs: 1.25M rows per GPU, D=1024, q=32, 30% missing) at 8
cd "$(dirname "$0")/.."
run() {  # n, tag, extra args
  n=$1; tag=$2; shift 2
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --steps 20 --warmup 3 --no-cpu "$@" > gpurun_out/scale_${tag}_${n}.txt 2> gpurun_out/scale_${tag}_${n}.err
  python - <<PY
import json
try:
    l = json.loads(open("gpurun_out/scale_${tag}_${n}.txt").read().strip().splitlines()[-1])
    print("${tag}", l["n_gpus"], "value %.4g" % l["value"], "ms/step %.3f" % l["ms_per_step"], "per-gpu %.4g" % (l["value"] / l["n_gpus"]), l["clocks"])
except Exception as e:
    print("${tag} ${n} FAILED", e); print(open("gpurun_out/scale_${tag}_${n}.err").read()[-1500:])
PY
}
python bench.py --steps 20 --warmup 3 --no-cpu --no-f32 > gpurun_out/scale_c2_1.txt 2> gpurun_out/scale_c2_1.err
python -c "
import json
l = json.loads(open('gpurun_out/scale_c2_1.txt').read().strip().splitlines()[-1]); print('c2 1 value %.4g ms/step %.3f' % (l['value'], l['ms_per_step']))"
for n in 2 4 8; do run $n c2; done
run 8 c3 --N 1250000 --D 1024 --q 32 --missing 0.3 --steps 5
