import sys; sys.path.insert(0, "/root/repo")
import numpy as np, torch
from pyvb_b200 import PlateEngine, LDSEngine
for (N, D, q, algo) in [(0, 32, 16, "auto"), (0, 5, 2, "generic"), (1, 32, 16, "auto"), (3, 64, 64, "auto"), (2, 32, 16, "auto")]:
    try:
        X = np.random.RandomState(0).randn(N, D)
        if N: X[0, 0] = np.nan
        for prec in ("f64", "f32"):
            if prec == "f32" and (q not in (16, 32) or N == 0 and False):
                continue
            e = PlateEngine(X, q, mode="B", algo=algo, precision=prec)
            e.init_random(seed=1)
            v = [e.iterate() for _ in range(2)]
            e.check()
            print("OK", N, D, q, algo, prec, v)
    except Exception as ex:
        print("FAIL", N, D, q, algo, repr(ex)[:300])
try:
    e = LDSEngine(np.zeros((0, 10, 3)), 2); e.iterate(); print("OK lds B=0")
except Exception as ex:
    print("FAIL lds B=0", repr(ex)[:200])
