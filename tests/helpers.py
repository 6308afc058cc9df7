"""Shared helpers for the parity tests (test infrastructure)."""
import os

import numpy as np

from oracle.plate_oracle import PlateOracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

STATE_KEYS = ("Wbar", "Wvar", "mu", "muvar", "Zbar", "Sig", "Xhat", "V", "qb")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def golden_state(g, prefix):
    st = {k: g[prefix + k] for k in STATE_KEYS}
    if prefix + "al_qb" in g:
        st["al_qb"] = g[prefix + "al_qb"]
    return st


def tensor_rel(a, b):
    """Tensor-wise relative error max|a-b| / max|b|."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = max(float(np.max(np.abs(b))), 1e-300)
    return float(np.max(np.abs(a - b))) / den


def oracle_from_golden(g, mode="A"):
    o = PlateOracle(g["X"], int(g["q"]), mode=mode, ard=bool(int(g.get("ard", 0))))
    o.load_state(golden_state(g, "init_"))
    return o
