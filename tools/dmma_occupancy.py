"""FP64 tensor (DMMA) throughput of the pure-DMMA loop vs resident warps per SM (8 independent accumulators per warp)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyvb_b200 import _cabi
lib = _cabi.lib(); dev = torch.device("cuda", 0)
st = torch.cuda.current_stream(dev).cuda_stream
for blocks_per_sm in (1, 2, 3, 4):
    blocks = 148 * blocks_per_sm
    scratch = torch.empty(blocks * 256, dtype=torch.float64, device=dev)
    iters = 20000
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); _cabi.check(lib.pyvb_bench_dmma_f64(blocks, iters, scratch.data_ptr(), st), "b"); e1.record()
        torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    print("warps/SM %2d: %.2f TF/s" % (8 * blocks_per_sm, blocks * 8 * iters * 8 * 512.0 / (best * 1e-3) * 1e-12), flush=True)
