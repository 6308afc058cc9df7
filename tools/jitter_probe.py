"""Per-sweep and per-phase device times on every rank, with and without the exchange (diagnosis)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from bench import make_data
from pyvb_b200 import PlateEngine

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N, D, q = 1000000, 256, 16
X = make_data(torch, N, D, q, 0.2, 1234 + rank, dev)
eng = PlateEngine(X, q, mode="B", keep_sigma=False, distributed=True, row_offset=rank * N, device=dev)
eng.init_random(seed=4321, rank=rank)
for mode in ["comm", "nocomm", "comm"]:
    eng.distributed = (mode == "comm")
    for _ in range(5):
        eng.iterate_async()
    dist.barrier(); torch.cuda.synchronize()
    K = 12
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(K)]
    for k in range(K):
        ev[k][0].record(); eng.update_W(); eng._ensure_gw()
        ev[k][1].record(); eng.update_Z()
        ev[k][2].record(); eng._ensure_stats()
        ev[k][3].record(); eng._global(1 | 2 | 4 | 8, eng.trace.data_ptr())
        ev[k][4].record()
    torch.cuda.synchronize()
    tot = [ev[k][0].elapsed_time(ev[k][4]) for k in range(K)]
    ph = [[ev[k][i].elapsed_time(ev[k][i + 1]) for i in range(4)] for k in range(K)]
    wall = ev[0][0].elapsed_time(ev[K - 1][4]) / K
    print("rank %d %-6s mean %.3f  sweeps %s" % (rank, mode, wall, " ".join("%.2f" % t for t in tot)), flush=True)
    print("rank %d %-6s phases W %.3f Z %.3f stats %.3f global %.3f" % (
        rank, mode, *[sum(p[i] for p in ph) / K for i in range(4)]), flush=True)
eng.close()
dist.destroy_process_group()
