"""tcgen05.mma rate microbenchmark: SM clocks per instruction for kind::i8 and kind::f16 (bf16) at M = 128, with the
operand buffers rotating, a commit per pair, and background bulk-copy traffic into shared memory."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyvb_b200 import _cabi
lib = _cabi.lib()
out = torch.zeros(296, dtype=torch.int64, device="cuda")
src = torch.ones(148 << 20, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
iters = 4000
for kind, name in ((0, "i8"), (1, "bf16")):
    for n in (128, 224, 256):
        for mode in (0, 1, 2, 3, 4, 7):
            out.zero_()
            _cabi.check(lib.pyvb_bench_umma(148, iters, n, kind, mode, src.data_ptr(), out.data_ptr(), st), "bench")
            torch.cuda.synchronize()
            c = out[:148].double().mean().item() / (2 * iters)
            fills = out[148:].double().mean().item()
            macs = 128 * n * (32 if kind == 0 else 16)
            print("%-4s N=%3d mode=%d (rotate %d commit %d fill %d): %6.1f clk / MMA -> %5.0f MAC/clk/SM   fill %.1f KB per MMA pair"
                  % (name, n, mode, mode & 1, (mode >> 1) & 1, (mode >> 2) & 1, c, macs / c, fills * 14.0 / iters), flush=True)
