"""Golden fixtures of the LDS VB smoother (BASELINE config 5) by RUNNING THE LITERAL REFERENCE.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/make_ref.py && python oracle/gen_golden_lds.py

The graph and the sweep are exactly examples/Linear_Dynamic_System.py:47-76: columns of A and C as Gaussians under
hstack, DiagonalGamma Q and R, X_0 ~ N(0, I), X_t ~ N(A X_{t-1}, Q), Y_t ~ N(C X_t, R) observed; per iteration all
X_t forwards, all X_t backwards, the A columns, the C columns, Q, R.  The random initial state is snapshotted
before the first iteration, the state after every iteration.

Fixtures:  lds_a.npz (q=2, d=5, T=30: the shipped shape, shorter), lds_b.npz (q=3, d=4, T=17), lds_c.npz (q=8, d=5, T=12)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_ref import import_ref, make_ref  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


def simulate(q, d, T, rng):
    """examples/Linear_Dynamic_System.py:20-44 with an explicit RandomState."""
    A = rng.randn(q, q)
    while np.max(np.abs(np.linalg.eig(A)[0])) > 1:
        A = rng.randn(q, q) * 0.7
    C = rng.randn(d, q) * 10
    Rc = np.linalg.cholesky(np.diag(rng.rand(d)) * 0.1)
    Qc = np.linalg.cholesky(np.diag(rng.rand(q)) * 0.1)
    X = np.zeros((T, q))
    Y = np.zeros((T, d))
    X[0] = rng.randn(q)
    Y[0] = C @ X[0] + Rc @ rng.randn(d)
    for t in range(1, T):
        X[t] = A @ X[t - 1] + Qc @ rng.randn(q)
        Y[t] = C @ X[t] + Rc @ rng.randn(d)
    return Y


def build(pyvb, Y, q):
    nodes = pyvb.nodes
    T, d = Y.shape
    As = [nodes.Gaussian(q, np.zeros((q, 1)), np.eye(q) * 1e-3) for i in range(q)]
    A = nodes.hstack(As)
    Cs = [nodes.Gaussian(d, np.zeros((d, 1)), np.eye(d) * 1e-3) for i in range(q)]
    C = nodes.hstack(Cs)
    Q = nodes.DiagonalGamma(q, np.ones(q) * 1e-3, np.ones(q) * 1e-3)
    R = nodes.DiagonalGamma(d, np.ones(d) * 1e-3, np.ones(d) * 1e-3)
    X0 = nodes.Gaussian(q, np.zeros((q, 1)), np.eye(q))
    Y0 = nodes.Gaussian(d, C * X0, R)
    Y0.observe(Y[0].reshape(d, 1))
    Xs, Ys = [X0], [Y0]
    for t in range(1, T):
        Xs.append(nodes.Gaussian(q, A * Xs[-1], Q))
        Ys.append(nodes.Gaussian(d, C * Xs[-1], R))
        Ys[-1].observe(Y[t].reshape(d, 1))
    return dict(As=As, Cs=Cs, Q=Q, R=R, Xs=Xs)


def snapshot(m, prefix, out):
    q = len(m["As"])
    out[prefix + "A"] = np.hstack([a.qmu for a in m["As"]])
    out[prefix + "Avar"] = np.stack([np.diag(a.qcov) for a in m["As"]], 1)
    out[prefix + "C"] = np.hstack([c.qmu for c in m["Cs"]])
    out[prefix + "Cvar"] = np.stack([np.diag(c.qcov) for c in m["Cs"]], 1)
    d = out[prefix + "C"].shape[0]
    out[prefix + "Qa"] = np.asarray(m["Q"].qa, dtype=np.float64) * np.ones(q)
    out[prefix + "Qb"] = np.asarray(m["Q"].qb, dtype=np.float64) * np.ones(q)
    out[prefix + "Ra"] = np.asarray(m["R"].qa, dtype=np.float64) * np.ones(d)
    out[prefix + "Rb"] = np.asarray(m["R"].qb, dtype=np.float64) * np.ones(d)
    out[prefix + "X"] = np.stack([x.qmu[:, 0] for x in m["Xs"]])
    out[prefix + "Xcov"] = np.stack([x.qcov for x in m["Xs"]])
    offd = 0.0
    for n in list(m["As"]) + list(m["Cs"]):
        c = n.qcov
        offd = max(offd, float(np.max(np.abs(c - np.diag(np.diag(c))))))
    out[prefix + "max_offdiag"] = np.float64(offd)


def run(pyvb, Y, q, seed, niters, known=None):
    np.random.seed(seed)
    m = build(pyvb, Y, q)
    if known is not None:              # examples/LDS_knowns_in_A.py:72-74: observe the known entries of A's columns
        for i, a in enumerate(m["As"]):
            if not np.all(np.isnan(known[:, i])):
                a.observe(known[:, i].reshape(q, 1))
    out = {"Y": Y, "q": np.int64(q), "niters": np.int64(niters)}
    if known is not None:
        out["A_known"] = known
    snapshot(m, "init_", out)
    Xs = m["Xs"]
    for it in range(niters):
        [x.update() for x in Xs]
        Xs.reverse()
        [x.update() for x in Xs]
        Xs.reverse()
        [a.update() for a in m["As"]]
        [c.update() for c in m["Cs"]]
        m["Q"].update()
        m["R"].update()
        snapshot(m, "it%d_" % it, out)
    return out


def main():
    make_ref(quiet=True)
    pyvb = import_ref()
    assert pyvb is not None, "reference not available"
    os.makedirs(GOLD, exist_ok=True)
    for name, (q, d, T), seed, niters in [("lds_a", (2, 5, 30), 0, 6), ("lds_b", (3, 4, 17), 1, 6), ("lds_c", (8, 5, 12), 2, 5)]:
        Y = simulate(q, d, T, np.random.RandomState(100 + seed))
        out = run(pyvb, Y, q, seed, niters)
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
        print(name, "max offdiag of column covariances", max(float(out["it%d_max_offdiag" % i]) for i in range(niters)),
              "Qb", out["it%d_Qb" % (niters - 1)][:3])


def main_known():
    """lds_known.npz: q = 2 with A[0][0] = 1 and A[0][1] = dt known (the shipped LDS_knowns_in_A.py pattern), and
    lds_known_b.npz: q = 3 with a known entry in every column, one column fully known but one entry."""
    make_ref(quiet=True)
    pyvb = import_ref()
    assert pyvb is not None, "reference not available"
    nan = np.nan
    for name, (q, d, T), seed, known in [
            ("lds_known", (2, 5, 25), 3, np.array([[1.0, 0.05], [nan, nan]])),
            ("lds_known_b", (3, 4, 14), 4, np.array([[0.9, nan, 0.1], [nan, 0.8, nan], [nan, 0.2, nan]]))]:
        Y = simulate(q, d, T, np.random.RandomState(200 + seed))
        out = run(pyvb, Y, q, seed, 6, known=known)
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
        print(name, "A after 6 iterations", out["it5_A"].round(4).tolist())


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "known":
        main_known()
    else:
        main()
