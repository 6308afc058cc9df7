cd /root/repo
for ns in "" 1; do for nc in "" 1; do
PYVB_NOSAMPLER=$ns PYVB_NOCOMM=$nc python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('nosampler=$ns nocomm=$nc', l['ms_per_step'], l['kernels']['zstep_ms'], l['kernels']['stats_ms'])"
done; done
