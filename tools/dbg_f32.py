import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
from pyvb_b200 import PlateEngine
from oracle.plate_oracle import PlateOracle, synth_pca
N, D, q = 3000, 256, 16
X = synth_pca(N, D, q, 0.25, seed=N)
rng = np.random.RandomState(11)
init = {"Wbar": rng.randn(D, q), "Wvar": np.ones((D, q)), "mu": np.zeros(D), "muvar": np.ones(D),
        "Zbar": rng.randn(N, q), "Sig": np.tile(np.eye(q), (N, 1, 1)), "qb": 0.5}
o = PlateOracle(X, q, mode="B"); o.load_state(init)
e = PlateEngine(X, q, mode="B", device="cuda:0", precision="f32"); e.set_state(init)
e64 = PlateEngine(X, q, mode="B", device="cuda:0"); e64.set_state(init)
for it in range(2):
    ref, got, g64 = o.iterate(), e.iterate(), e64.iterate()
    print(it, ref, got, g64)
    print(" gl f32", e.gl[:12].cpu().numpy())
    print(" gl f64", e64.gl[:12].cpu().numpy())
    L = e.L
    v32, v64 = L.views(e.stats.cpu().numpy()), L.views(e64.stats.cpu().numpy())
    for k in v32:
        a, b = v32[k], v64[k]
        print("  ", k, "finite", np.isfinite(a).all(), "rel", np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
    print("   scal32", v32["scal"][:10]); print("   scal64", v64["scal"][:10])
