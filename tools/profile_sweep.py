"""One steady-state sweep of the default (INT8) path between cudaProfilerStart/Stop, for
    ncu --profile-from-start off --set full ... python tools/profile_sweep.py N D q missing
(every kernel of the sweep is captured exactly once)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from pyvb_b200 import PlateEngine  # noqa: E402
from tools.time_i8 import synth  # noqa: E402

N, D, q, missing = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])
dev = torch.device("cuda:0")
X = synth(N, D, q, missing, dev)
e = PlateEngine(X, q, mode="B", keep_sigma=False)
e.init_random(seed=5)
for _ in range(3):
    e.iterate_async()
torch.cuda.synchronize()
torch.cuda.profiler.start()
e.iterate_async()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
e.check()
print("use_i8", e.use_i8, "i8 stats", e.use_i8_stats, "elbo", float(e.trace[e.trace_pos - 1 if e.trace_pos else 0].item()))
