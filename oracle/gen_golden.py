"""Generate the golden fixtures under tests/golden/ by RUNNING THE LITERAL REFERENCE.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/make_ref.py && python oracle/gen_golden.py

The reference holds no known-answer vectors of its own, so
every fixture is: seed -> build the graph exactly as examples/PCA_missing_data.py:31-42
does -> snapshot the random initial state -> run k sweeps with the reference's
own ``Network.learn`` (stop rule disabled) -> dump the state after every sweep.

Fixtures:
  c1_shipped.npz     the shipped workload, np.random.seed(0), 25 sweeps
  small_a.npz        N=60,D=7,q=3, 30% missing, partial row 0, one all-NaN row
  small_b.npz        N=40,D=12,q=4, 50% missing, all-NaN row 0
  full_manual.npz    fully observed, src/tests.py:312-316 manual order (W, Mu, Z, Beta)
  ard.npz            Gamma precision per W column (ARD), network order
  modeB_ops.npz      generic operators with per-row Constant(tau*diag(mask_n)) precision
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_ref import import_ref, make_ref  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


def build(pyvb, X, q, ard=False, row_prec=None):
    """examples/PCA_missing_data.py:31-42 (ard: Gamma precision per column)."""
    nodes = pyvb.nodes
    N, d = X.shape
    if ard:
        Alphas = [nodes.Gamma(d, 1e-3, 1e-3) for i in range(q)]
        Ws = [nodes.Gaussian(d, np.zeros((d, 1)), Alphas[i]) for i in range(q)]
    else:
        Alphas = []
        Ws = [nodes.Gaussian(d, np.zeros((d, 1)), np.eye(d) * 1e-3) for i in range(q)]
    W = nodes.hstack(Ws)
    Mu = nodes.Gaussian(d, np.zeros((d, 1)), np.eye(d) * 1e-3)
    Beta = nodes.Gamma(d, 1e-3, 1e-3)
    Zs = [nodes.Gaussian(q, np.zeros((q, 1)), np.eye(q)) for i in range(N)]
    if row_prec is None:
        Xs = [nodes.Gaussian(d, W * z + Mu, Beta) for z in Zs]
    else:
        Xs = [nodes.Gaussian(d, W * z + Mu, row_prec[n]) for n, z in enumerate(Zs)]
    [xnode.observe(xval.reshape(d, 1)) for xnode, xval in zip(Xs, X)]
    return dict(Ws=Ws, W=W, Mu=Mu, Beta=Beta, Zs=Zs, Xs=Xs, Alphas=Alphas)


def snapshot(m, prefix, out):
    Ws, Mu, Beta, Zs, Xs, Alphas = m["Ws"], m["Mu"], m["Beta"], m["Zs"], m["Xs"], m["Alphas"]
    out[prefix + "Wbar"] = np.hstack([w.qmu for w in Ws])
    out[prefix + "Wvar"] = np.stack([np.diag(w.qcov) for w in Ws], 1)
    out[prefix + "mu"] = Mu.qmu[:, 0].copy()
    out[prefix + "muvar"] = np.diag(Mu.qcov).copy()
    out[prefix + "Zbar"] = np.stack([z.qmu[:, 0] for z in Zs])
    out[prefix + "Sig"] = np.stack([z.qcov for z in Zs])
    out[prefix + "Xhat"] = np.stack([x.qmu[:, 0] for x in Xs])
    out[prefix + "V"] = np.stack([np.diag(x.qcov) for x in Xs])
    out[prefix + "qb"] = np.float64(Beta.qb)
    out[prefix + "qa"] = np.float64(Beta.qa)
    if Alphas:
        out[prefix + "al_qb"] = np.array([a.qb for a in Alphas], dtype=np.float64)
        out[prefix + "al_qa"] = np.array([a.qa for a in Alphas], dtype=np.float64)
    offd = 0.0
    for n in list(Ws) + [Mu] + list(Xs):
        c = n.qcov
        offd = max(offd, float(np.max(np.abs(c - np.diag(np.diag(c))))))
    out[prefix + "max_offdiag"] = np.float64(offd)


def run_network(pyvb, X, q, seed, niters, ard=False):
    if seed is not None:                  # None: keep drawing from the current global stream
        np.random.seed(seed)
    m = build(pyvb, X, q, ard=ard)
    net = pyvb.Network()
    net.addnode(m["W"])
    net.fetch_network()
    out = {"X": X, "q": np.int64(q), "ard": np.int64(ard), "niters": np.int64(niters)}
    snapshot(m, "init_", out)
    net.find_iterable()
    order = []
    for n in net.iterable_nodes:
        for key in ("Ws", "Zs", "Xs", "Alphas"):
            if any(n is e for e in m[key]):
                order.append(key[0] if key != "Alphas" else "L")
                break
        else:
            order.append("M" if n is m["Mu"] else "B")
    out["order"] = np.array("".join(order))
    elbo = []
    for it in range(niters):
        net.learn(1, tol=-np.inf)         # the reference's own sweep + ELBO sum (network.py:46-49)
        elbo.append(net.llb)
        snapshot(m, "it%d_" % it, out)
    out["elbo"] = np.array(elbo, dtype=np.float64)
    return out


def run_manual(pyvb, X, q, seed, niters):
    """src/tests.py:312-316: explicit order W cols, Mu, Z rows, Beta (fully observed data)."""
    np.random.seed(seed)
    m = build(pyvb, X, q)
    out = {"X": X, "q": np.int64(q), "ard": np.int64(0), "niters": np.int64(niters)}
    snapshot(m, "init_", out)
    net = pyvb.Network()
    net.addnode(m["W"])
    net.fetch_network()
    net.find_iterable()
    elbo = []
    for it in range(niters):
        [w.update() for w in m["Ws"]]
        m["Mu"].update()
        [z.update() for z in m["Zs"]]
        m["Beta"].update()
        if it == 0:
            # X nodes are observed (no q_ln_det needed); every latent node has been updated once
            pass
        elbo.append(float(np.sum([n.log_lower_bound() for n in net.iterable_nodes])))
        snapshot(m, "it%d_" % it, out)
    out["elbo"] = np.array(elbo, dtype=np.float64)
    return out


def run_modeB_ops(pyvb, X, q, seed, niters, tau):
    """Mode B through the reference's *generic* operators: X_n observed (NaN -> 0)
    with precision parent Constant(tau*diag(mask_n)); sweeps of W cols, Z rows, Mu
    (node.py:203-227, nodes_todo.py:43-62, node.py:105-109).  Beta is not a node here."""
    np.random.seed(seed)
    mask = ~np.isnan(X)
    X0 = np.where(mask, X, 0.0)
    row_prec = [np.diag(tau * mask[n].astype(np.float64)) for n in range(X.shape[0])]
    m = build(pyvb, X0, q, row_prec=row_prec)
    out = {"X": X, "q": np.int64(q), "tau": np.float64(tau), "niters": np.int64(niters)}
    snapshot(m, "init_", out)
    for it in range(niters):
        [w.update() for w in m["Ws"]]
        [z.update() for z in m["Zs"]]
        m["Mu"].update()
        snapshot(m, "it%d_" % it, out)
        out["it%d_qldZ" % it] = np.array([z.q_ln_det for z in m["Zs"]], dtype=np.float64)
        out["it%d_qldW" % it] = np.array([w.q_ln_det for w in m["Ws"]], dtype=np.float64)
        out["it%d_qldMu" % it] = np.float64(m["Mu"].q_ln_det)
    return out


def shipped_data(seed=0):
    """examples/PCA_missing_data.py:11-27 verbatim, with the seed set first."""
    np.random.seed(seed)
    q, d, N, Nmissing = 2, 5, 200, 100
    true_W = np.random.randn(d, q)
    true_Z = np.random.randn(N, q)
    true_mean = np.random.randn(d, 1)
    true_prec = 20.
    Xdata_full = np.dot(true_Z, true_W.T) + true_mean.T
    Xdata_observed = Xdata_full + np.random.randn(N, d) * np.sqrt(1. / true_prec)
    missing_index_i = np.argsort(np.random.randn(N))[:Nmissing]
    missing_index_j = np.random.multinomial(1, np.ones(d) / d, Nmissing).nonzero()[1]
    Xdata = Xdata_observed.copy()
    Xdata[missing_index_i, missing_index_j] = np.nan
    return Xdata, q


def synth(N, D, q, missing, seed):
    rng = np.random.RandomState(seed)
    W = rng.randn(D, q)
    Z = rng.randn(N, q)
    mu = rng.randn(D)
    X = Z @ W.T + mu[None, :] + rng.randn(N, D) * np.sqrt(1.0 / 20.0)
    if missing > 0:
        X[rng.rand(N, D) < missing] = np.nan
    return X


def main():
    warnings.simplefilter("ignore")
    make_ref(quiet=True)
    pyvb = import_ref()
    if pyvb is None:
        raise SystemExit("translated reference not available (need /root/reference)")
    os.makedirs(GOLD, exist_ok=True)

    # (i) shipped workload: np.random.seed(0) inserted before examples/PCA_missing_data.py:16;
    # data and the nodes' random init come from ONE global stream, exactly as in the script.
    X, q = shipped_data(0)
    np.savez_compressed(os.path.join(GOLD, "c1_shipped.npz"), **run_network(pyvb, X, q, seed=None, niters=25))

    # (ii) small shapes with a partial row 0 / an all-NaN row
    X = synth(60, 7, 3, 0.30, seed=11)
    X[0, :] = synth(1, 7, 3, 0.0, seed=12)[0]
    X[0, 2] = np.nan                      # partial row 0
    X[17, :] = np.nan                     # latent row
    np.savez_compressed(os.path.join(GOLD, "small_a.npz"), **run_network(pyvb, X, 3, seed=2, niters=12))
    X = synth(40, 12, 4, 0.50, seed=21)
    X[0, :] = np.nan                      # latent row 0
    np.savez_compressed(os.path.join(GOLD, "small_b.npz"), **run_network(pyvb, X, 4, seed=3, niters=12))

    # (iii) fully observed, manual order
    X = synth(50, 6, 2, 0.0, seed=31)
    np.savez_compressed(os.path.join(GOLD, "full_manual.npz"), **run_manual(pyvb, X, 2, seed=4, niters=10))

    # (iv) ARD
    X = synth(48, 8, 3, 0.25, seed=41)
    np.savez_compressed(os.path.join(GOLD, "ard.npz"), **run_network(pyvb, X, 3, seed=5, niters=10, ard=True))

    # (v) mode B via generic operators
    X = synth(36, 9, 3, 0.35, seed=51)
    X[5, :] = np.nan
    np.savez_compressed(os.path.join(GOLD, "modeB_ops.npz"), **run_modeB_ops(pyvb, X, 3, seed=6, niters=4, tau=7.5))
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
