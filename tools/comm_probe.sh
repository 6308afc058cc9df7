# exchange cost at N GPUs (default 2): fused peer-memory exchange vs NCCL all_reduce vs no exchange, C2 per GPU
cd "$(dirname "$0")/.."
n=${1:-2}
for mode in peer nccl none; do
nc=""; [ $mode = none ] && nc=1
PYVB_COMM=$mode PYVB_NOCOMM=$nc python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 3 --no-cpu --no-f32 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('gpus=$n exchange=$mode ms/sweep %.4f' % l['ms_per_step'])"
done
