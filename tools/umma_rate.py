"""tcgen05.mma issue-rate microbenchmark: SM clocks per instruction for kind::i8 and kind::f16 (bf16) at M = 128."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyvb_b200 import _cabi
lib = _cabi.lib()
out = torch.zeros(148, dtype=torch.int64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for blocks in (1, 148):
    for kind, name in ((0, "i8"), (1, "bf16")):
        for n in (64, 128, 224, 256):
            for iters in (2000,):
                _cabi.check(lib.pyvb_bench_umma(blocks, iters, n, kind, out.data_ptr(), st), "bench")
                torch.cuda.synchronize()
                c = out[:blocks].double().mean().item() / (2 * iters)
                macs = 128 * n * (32 if kind == 0 else 16)
                print("blocks %3d %-4s N=%3d: %.1f clk / MMA  -> %.0f MAC/clk/SM" % (blocks, name, n, c, macs / c), flush=True)
