"""GPU parity tests: the CUDA path (through the C-ABI) against the golden vectors produced by the
literal reference and against the CPU oracle.  FP64 tolerance: 1e-9 tensor-wise relative on every
state tensor and 1e-9 relative on the ELBO, iteration by iteration (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from helpers import golden_state, load_golden, tensor_rel
from oracle.plate_oracle import PlateOracle, synth_pca

pytestmark = pytest.mark.gpu

TOL = 1e-9


@pytest.fixture(scope="module")
def eng():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from pyvb_b200 import PlateEngine
    return PlateEngine


def _cmp_state(st, ref, keys, tag):
    for k in keys:
        r = tensor_rel(st[k], ref[k])
        assert r < TOL, (tag, k, r)


def rand_init(N, D, q, seed=0):
    rng = np.random.RandomState(seed)
    return {"Wbar": rng.randn(D, q), "Wvar": rng.rand(D, q) + 0.5, "mu": rng.randn(D) * 0.1,
            "muvar": np.ones(D), "Zbar": rng.randn(N, q),
            "Sig": np.tile(np.eye(q), (N, 1, 1)) * (rng.rand(N, 1, 1) + 0.5), "qb": 0.7}


@pytest.mark.parametrize("name", ["c1_shipped.npz", "small_a.npz", "small_b.npz", "ard.npz"])
def test_modeA_matches_literal_reference(eng, name):
    g = load_golden(name)
    q = int(g["q"])
    e = eng(g["X"], q, mode="A", ard=bool(int(g["ard"])), algo="generic")
    e.set_state(golden_state(g, "init_"))
    for it in range(int(g["niters"])):
        elbo = e.iterate()
        st = e.get_state()
        ref = golden_state(g, "it%d_" % it)
        _cmp_state(st, ref, ("Wbar", "Wvar", "mu", "muvar", "Zbar", "Sig", "Xhat", "V"), (name, it))
        assert abs(st["qb"] - float(ref["qb"])) <= TOL * abs(float(ref["qb"]))
        if "al_qb" in ref:
            assert tensor_rel(st["al_qb"], ref["al_qb"]) < TOL
        assert abs(elbo - g["elbo"][it]) <= TOL * abs(g["elbo"][it]), (name, it, elbo, g["elbo"][it])
    e.check()


def test_modeB_updates_match_generic_operators(eng):
    g = load_golden("modeB_ops.npz")
    q = int(g["q"])
    e = eng(g["X"], q, mode="B", algo="generic")
    st0 = golden_state(g, "init_")
    st0["qb"] = e.qa / float(g["tau"])          # fix tau = qa/qb
    e.set_state(st0)
    for it in range(int(g["niters"])):
        e.update_W(); e.update_Z(); e.update_Mu()
        st = e.get_state()
        ref = {k: g["it%d_%s" % (it, k)] for k in ("Wbar", "Wvar", "mu", "muvar", "Zbar", "Sig")}
        _cmp_state(st, ref, ref.keys(), ("modeB_ops", it))
        ld = e.logdet.cpu().numpy()
        qld = g["it%d_qldZ" % it]
        fin = np.isfinite(qld)
        assert tensor_rel(0.5 / ld[fin], qld[fin]) < TOL
        assert np.all(ld[~fin] == 0.0)            # all-NaN row: posterior = prior, log prod diag chol = 0


@pytest.mark.parametrize("shape", [(1, 3, 1), (37, 5, 2), (513, 33, 7), (2000, 48, 8), (300, 20, 33), (130, 9, 64)])
@pytest.mark.parametrize("mode", ["A", "B"])
def test_generic_vs_oracle(eng, shape, mode):
    N, D, q = shape
    X = synth_pca(N, D, q, 0.3, seed=N)
    if N > 40:
        X[3, :] = np.nan
    init = rand_init(N, D, q, seed=1)
    if mode == "A":
        rng = np.random.RandomState(5)
        init["Xhat"] = np.where(np.isnan(X), rng.randn(N, D), X)
        init["V"] = np.where(np.isnan(X), 1.3, 0.0)
    o = PlateOracle(X, q, mode=mode)
    o.load_state(init)
    e = eng(X, q, mode=mode, algo="generic")
    e.set_state(init)
    for it in range(4):
        ref = o.iterate()
        got = e.iterate()
        st = e.get_state()
        _cmp_state(st, o.state(), ("Wbar", "Wvar", "mu", "muvar", "Zbar", "Sig"), (shape, mode, it))
        assert abs(st["qb"] - o.qb) <= TOL * abs(o.qb)
        if np.isfinite(ref):
            assert abs(got - ref) <= TOL * abs(ref), (shape, mode, it, got, ref)
        else:
            assert got == ref
    e.check()


@pytest.mark.parametrize("shape,algo", [((400, 24, 5), "generic"), ((600, 64, 64), "dmma"), ((900, 128, 16), "dmma")])
def test_ard_modeB_vs_oracle(eng, shape, algo):
    """ARD Gamma precisions per latent column (BASELINE config 4 is this at D = 512, q = 64): Gamma(d, a0, b0) as the
    precision parent of every W column, /root/reference/src/pyvb/nodes/nodes_todo.py:113-138."""
    N, D, q = shape
    X = synth_pca(N, D, q, 0.25, seed=9)
    init = rand_init(N, D, q, seed=2)
    init["al_qb"] = np.linspace(0.5, 1.5, q)
    o = PlateOracle(X, q, mode="B", ard=True)
    o.load_state(init)
    e = eng(X, q, mode="B", ard=True, algo=algo)
    e.set_state(init)
    for it in range(5):
        ref, got = o.iterate(), e.iterate()
        st = e.get_state()
        _cmp_state(st, o.state(), ("Wbar", "Wvar", "mu", "Zbar", "Sig"), ("ard", it))
        assert tensor_rel(st["al_qb"], o.al_qb) < TOL
        assert abs(got - ref) <= TOL * abs(ref)


def test_modeB_equals_modeA_when_nothing_missing(eng):
    N, D, q = 150, 12, 3
    X = synth_pca(N, D, q, 0.0, seed=4)
    init = rand_init(N, D, q, seed=3)
    ea, eb = eng(X, q, mode="A", algo="generic"), eng(X, q, mode="B", algo="generic")
    ea.set_state(init); eb.set_state(init)
    for it in range(4):
        a, b = ea.iterate(), eb.iterate()
        assert abs(a - b) <= 1e-12 * abs(a)
    sa, sb = ea.get_state(), eb.get_state()
    _cmp_state(sa, sb, ("Wbar", "mu", "Zbar", "Sig"), "AB")


def test_nonpd_raises_linalgerror(eng):
    X = synth_pca(50, 6, 2, 0.1, seed=1)
    e = eng(X, 2, mode="B", P0=-1e9 * np.eye(2), algo="generic")
    e.set_state(rand_init(50, 6, 2))
    e.update_Z()
    with pytest.raises(np.linalg.LinAlgError):
        e.check()


def test_stats_are_additive_over_row_shards(eng):
    # the quantity that is all-reduced: stats(rows A ++ rows B) == stats(A) + stats(B)
    N, D, q = 900, 16, 4
    X = synth_pca(N, D, q, 0.3, seed=2)
    init = rand_init(N, D, q)
    full = eng(X, q, mode="B", algo="generic"); full.set_state(init); full._ensure_stats()
    acc = None
    for lo, hi in [(0, 311), (311, 900)]:
        sub = {k: (v[lo:hi] if k in ("Zbar", "Sig") else v) for k, v in init.items()}
        p = eng(X[lo:hi], q, mode="B", algo="generic"); p.set_state(sub); p._ensure_stats()
        acc = p.stats.clone() if acc is None else acc + p.stats
    assert tensor_rel(acc.cpu().numpy(), full.stats.cpu().numpy()) < 1e-12


# ---------------------------------------------------------------- DMMA (FP64 tensor core) path
DMMA_SHAPES = [(5000, 256, 16), (64, 16, 16), (1, 32, 16), (777, 64, 32), (2100, 128, 32), (1000, 48, 8), (4099, 80, 8),
               (333, 48, 64), (1500, 128, 64)]


@pytest.mark.parametrize("shape", DMMA_SHAPES)
def test_dmma_zstep_and_stats_match_generic(eng, shape):
    """kernel-level: same inputs through the DMMA kernels and the generic kernels."""
    N, D, q = shape
    X = synth_pca(N, D, q, 0.3, seed=N + q)
    if N > 10:
        X[2, :] = np.nan            # an all-missing row
        X[5, :] = 1.0               # a fully observed row
    init = rand_init(N, D, q, seed=7)
    eg, ed = eng(X, q, mode="B", algo="generic"), eng(X, q, mode="B", algo="dmma")
    for e in (eg, ed):
        e.set_state(init)
        e.update_Z()
        e._ensure_stats()
    sg, sd = eg.get_state(), ed.get_state()
    for k in ("Zbar", "Sig"):
        assert tensor_rel(sd[k], sg[k]) < 1e-11, (shape, k)
    assert tensor_rel(ed.M2.contiguous().cpu().numpy(), eg.M2.contiguous().cpu().numpy()) < 1e-11
    lg, ld = eg.logdet.cpu().numpy(), ed.logdet.cpu().numpy()
    assert np.max(np.abs(lg - ld)) < 1e-11 * max(1.0, np.max(np.abs(lg)))
    vg, vd = eg.L.views(eg.stats.cpu().numpy()), ed.L.views(ed.stats.cpu().numpy())
    for k in ("T1", "Bst", "Ast", "cnt", "colx", "S", "zsum"):
        assert tensor_rel(vd[k], vg[k]) < 1e-11, (shape, k)
    fin = np.isfinite(vg["scal"])
    assert np.array_equal(fin, np.isfinite(vd["scal"]))
    assert tensor_rel(vd["scal"][fin], vg["scal"][fin]) < 1e-11
    ed.check()


@pytest.mark.parametrize("shape", [(3000, 256, 16), (1200, 64, 32), (900, 32, 8), (700, 96, 64)])
def test_dmma_iterations_match_oracle(eng, shape):
    N, D, q = shape
    X = synth_pca(N, D, q, 0.25, seed=N)
    init = rand_init(N, D, q, seed=11)
    o = PlateOracle(X, q, mode="B")
    o.load_state(init)
    e = eng(X, q, mode="B", algo="dmma")
    e.set_state(init)
    for it in range(5):
        ref, got = o.iterate(), e.iterate()
        st = e.get_state()
        _cmp_state(st, o.state(), ("Wbar", "Wvar", "mu", "muvar", "Zbar", "Sig"), (shape, it))
        assert abs(st["qb"] - o.qb) <= TOL * abs(o.qb)
        assert abs(got - ref) <= TOL * abs(ref), (shape, it, got, ref)
    e.check()


def test_dmma_rejects_unsupported_shape(eng):
    from pyvb_b200._cabi import PyvbError
    X = synth_pca(40, 10, 3, 0.1, seed=1)
    e = eng(X, 3, mode="B", algo="dmma")
    e.set_state(rand_init(40, 10, 3))
    with pytest.raises(PyvbError):
        e.update_Z()


def test_iterate_from_host_equals_resident_sweep(eng):
    """The end-to-end entry point (data shard uploaded from pinned host memory in chunks, Z step overlapped with
    the upload) must give exactly the sweep of the resident path."""
    import torch
    N, D, q = 5003, 64, 16
    X = synth_pca(N, D, q, 0.25, seed=9)
    init = rand_init(N, D, q, seed=4)
    a, b = eng(X, q, mode="B"), eng(X, q, mode="B")
    Xh = torch.as_tensor(X).pin_memory()
    for e in (a, b):
        e.set_state(init)
    b._ensure_stats()                           # statistics of the initial state (what the first W update uses) ...
    b.X.fill_(0.0)                              # ... from here on b only sees the data through the upload
    for it in range(3):
        ra = a.iterate()
        slot = b.iterate_from_host(Xh, nchunks=5)
        rb = float(b.trace[slot].item())
        assert abs(ra - rb) <= 1e-12 * abs(ra), (it, ra, rb)
    sa, sb = a.get_state(), b.get_state()
    for k in ("Wbar", "mu", "Zbar", "Sig"):
        assert tensor_rel(sb[k], sa[k]) < 1e-12, k


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_empty_row_block_sweeps_on_the_priors(eng, precision):
    """An empty shard (N = 0 rows) is legal: the sweep runs on the priors alone and stays finite (it still takes
    part in the exchange when rows are sharded)."""
    D, q = 32, 16
    e = eng(np.zeros((0, D)), q, mode="B", precision=precision)
    e.init_random(seed=3)
    vals = [e.iterate() for _ in range(3)]
    assert np.all(np.isfinite(vals))
    st = e.get_state()
    assert np.all(np.isfinite(st["Wbar"])) and np.allclose(st["Wbar"], 0.0)      # no data: W at its prior mean
    e.check()
