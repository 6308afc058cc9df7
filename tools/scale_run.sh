#!/bin/bash
# multi-GPU scaling runs on one box: C3 shard (N=1.25M rows per GPU, D=1024, q=32, 30% missing) at 2/4/8 GPUs, C2 at 8
cd "$(dirname "$0")/.."
run() {  # n, tag, extra args
  n=$1; tag=$2; shift 2
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --steps 5 --warmup 3 --no-cpu "$@" > gpurun_out/scale_${tag}_${n}.txt 2> gpurun_out/scale_${tag}_${n}.err
  tail -c 400 gpurun_out/scale_${tag}_${n}.txt | head -c 300; echo
}
for n in 2 4 8; do run $n c3 --N 1250000 --D 1024 --q 32 --missing 0.3; done
run 8 c2
