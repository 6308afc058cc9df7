"""Golden fixtures for the node-API rows of the drop-in by RUNNING THE LITERAL REFERENCE.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/make_ref.py && python oracle/gen_golden_api.py

Fixtures:
  simple_pca.npz   src/tests.py:176-202 (simple_PCA): q = 1, W is ONE Gaussian column (no hstack), products are
                   Multiplication(W, z_n) with scalar z_n; manual order W, Z rows, Mu, noise; 10 iterations
  simple_regression.npz  src/tests.py:100-128: y_n ~ N(x_n * A + B, noise) with constant scalar regressors; A, B, noise x 10
  messages.npz     the per-node messages of the shipped graph (examples/PCA_missing_data.py:31-37, N = 6, d = 4, q = 2,
                   one NaN) at its random initial state: pass_up_m1_m2 of X_n, Addition, Multiplication (to W: the
                   4-tuple, to z), hstack (to every column), and pass_down_ExxT of Addition / Multiplication
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_ref import import_ref, make_ref  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


def build_simple_pca(nodes, X):
    """src/tests.py:187-193"""
    N, d = X.shape
    noise = nodes.Gamma(d, 1e-3, 1e-3)
    W = nodes.Gaussian(d, np.zeros((d, 1)), np.eye(d) * 0.001)
    Mu = nodes.Gaussian(d, np.zeros((d, 1)), np.eye(d) * 0.001)
    Zs = [nodes.Gaussian(1, np.zeros((1, 1)), np.eye(1)) for i in range(N)]
    mults = [nodes.Multiplication(W, z) for z in Zs]
    Xs = [nodes.Gaussian(d, m + Mu, noise) for m in mults]
    [n.observe(v.reshape(d, 1)) for n, v in zip(Xs, X)]
    return noise, W, Mu, Zs, Xs


def snap_simple(prefix, out, noise, W, Mu, Zs):
    out[prefix + "W"] = W.qmu[:, 0].copy()
    out[prefix + "Wcov"] = np.array(W.qcov, copy=True)
    out[prefix + "mu"] = Mu.qmu[:, 0].copy()
    out[prefix + "mucov"] = np.array(Mu.qcov, copy=True)
    out[prefix + "Z"] = np.array([float(z.qmu) for z in Zs])
    out[prefix + "Zvar"] = np.array([float(z.qcov) for z in Zs])
    out[prefix + "qb"] = np.float64(noise.qb)
    out[prefix + "qa"] = np.float64(noise.qa)


def simple_pca(pyvb, seed=7, niters=10):
    np.random.seed(seed)
    N, d = 100, 8
    X = np.dot(np.random.randn(N, 1), np.random.randn(d, 1).T) + np.random.randn(d, 1).T + np.random.randn(N, d) * 0.1
    noise, W, Mu, Zs, Xs = build_simple_pca(pyvb.nodes, X)
    out = {"X": X, "niters": np.int64(niters), "seed": np.int64(seed)}
    snap_simple("init_", out, noise, W, Mu, Zs)
    for it in range(niters):
        W.update()
        [e.update() for e in Zs]
        Mu.update()
        noise.update()
        snap_simple("it%d_" % it, out, noise, W, Mu, Zs)
    return out


def build_regression(nodes, x, y):
    """src/tests.py:108-123"""
    N = x.shape[0]
    B = nodes.Gaussian(1, np.array([[0.]]), np.array([[1e-2]]))
    A = nodes.Gaussian(1, np.array([[0.]]), np.array([[1e-2]]))
    noise = nodes.Gamma(1, 1e-3, 1e-3)
    Xs = [nodes.Constant(xx.reshape(1, 1)) for xx in x]
    Ys = [nodes.Gaussian(1, Xnode * A + B, noise) for Xnode in Xs]
    for n, yy in zip(Ys, y):
        n.observe(yy.reshape(1, 1))
    return A, B, noise, Ys


def simple_regression(pyvb, seed=11, niters=10):
    np.random.seed(seed)
    N = 200
    x = np.linspace(-1, 1, N).reshape(N, 1)
    y = 0.7 * x + 0.3 + np.random.randn(N, 1) * np.sqrt(1. / 10.)
    A, B, noise, Ys = build_regression(pyvb.nodes, x, y)
    out = {"x": x, "y": y, "niters": np.int64(niters), "seed": np.int64(seed),
           "init": np.array([float(A.qmu[0, 0]), float(A.qcov[0, 0]), float(B.qmu[0, 0]), float(B.qcov[0, 0]), float(noise.qb)])}
    for it in range(niters):
        A.update()
        B.update()
        noise.update()
        out["it%d" % it] = np.array([float(A.qmu[0, 0]), float(A.qcov[0, 0]), float(B.qmu[0, 0]), float(B.qcov[0, 0]),
                                     float(noise.qb), float(noise.pass_down_Ex()[0, 0])])
    return out


def build_shipped(nodes, X, q):
    N, d = X.shape
    Ws = [nodes.Gaussian(d, np.zeros((d, 1)), np.eye(d) * 1e-3) for i in range(q)]
    W = nodes.hstack(Ws)
    Mu = nodes.Gaussian(d, np.zeros((d, 1)), np.eye(d) * 1e-3)
    Beta = nodes.Gamma(d, 1e-3, 1e-3)
    Zs = [nodes.Gaussian(q, np.zeros((q, 1)), np.eye(q)) for i in range(N)]
    Xs = [nodes.Gaussian(d, W * z + Mu, Beta) for z in Zs]
    [xn.observe(xv.reshape(d, 1)) for xn, xv in zip(Xs, X)]
    return Ws, W, Mu, Beta, Zs, Xs


def messages(nodes_mod, seed=3):
    """Every message of the shipped graph at its random initial state (no update has run)."""
    rng = np.random.RandomState(50)
    N, d, q = 6, 4, 2
    X = rng.randn(N, d)
    X[2, 1] = np.nan
    np.random.seed(seed)
    Ws, W, Mu, Beta, Zs, Xs = build_shipped(nodes_mod, X, q)
    out = {"X": X, "q": np.int64(q), "seed": np.int64(seed)}
    for n, x in enumerate(Xs):
        add = x.mean_parent
        mult = add.A
        m = x.pass_up_m1_m2(add)
        out["x%d_m1" % n], out["x%d_m2" % n] = m
        m = add.pass_up_m1_m2(mult)
        out["add%d_to_mult_m1" % n], out["add%d_to_mult_m2" % n] = m
        m = add.pass_up_m1_m2(Mu)
        out["add%d_to_mu_m1" % n], out["add%d_to_mu_m2" % n] = m
        m = mult.pass_up_m1_m2(Zs[n])
        out["mult%d_to_z_m1" % n], out["mult%d_to_z_m2" % n] = m
        m = mult.pass_up_m1_m2(W)
        for k, v in enumerate(m):
            out["mult%d_to_W_%d" % (n, k)] = v
        out["add%d_ExxT" % n] = add.pass_down_ExxT()
        out["mult%d_ExxT" % n] = mult.pass_down_ExxT()
        out["mult%d_Ex" % n] = mult.pass_down_Ex()
    for i, w in enumerate(Ws):
        m = W.pass_up_m1_m2(w)
        out["hstack_to_w%d_m1" % i], out["hstack_to_w%d_m2" % i] = m
    out["W_ExxT"] = W.pass_down_ExxT()
    return {k: np.asarray(v, dtype=np.float64) for k, v in out.items()}


def main():
    warnings.simplefilter("ignore")
    make_ref(quiet=True)
    pyvb = import_ref()
    if pyvb is None:
        raise SystemExit("translated reference not available (need /root/reference)")
    os.makedirs(GOLD, exist_ok=True)
    np.savez_compressed(os.path.join(GOLD, "simple_pca.npz"), **simple_pca(pyvb))
    np.savez_compressed(os.path.join(GOLD, "messages.npz"), **messages(pyvb.nodes))
    np.savez_compressed(os.path.join(GOLD, "simple_regression.npz"), **simple_regression(pyvb))
    for f in ("simple_pca.npz", "messages.npz", "simple_regression.npz"):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
