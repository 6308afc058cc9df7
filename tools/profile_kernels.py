"""Small driver for ncu: a few sweeps of the FP64 (DMMA) or FP32 (tcgen05) path on a C2- or C3-shaped shard."""
import argparse, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_data
from pyvb_b200 import PlateEngine

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=296 * 64 * 8)
ap.add_argument("--D", type=int, default=256)
ap.add_argument("--q", type=int, default=16)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--precision", default="f64")
a = ap.parse_args()
dev = torch.device("cuda", 0)
X = make_data(torch, a.N, a.D, a.q, 0.2, 1234, dev)
e = PlateEngine(X, a.q, mode="B", algo="auto", keep_sigma=False, device=dev, precision=a.precision)
e.init_random(seed=4321)
for _ in range(a.reps):
    e.iterate_async()
torch.cuda.synchronize()
e.check()
print("ok", a.precision, float(e.trace[a.reps - 1]))
