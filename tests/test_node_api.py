"""The node-level API of the drop-in against the LITERAL reference (fixtures of oracle/gen_golden_api.py).

CPU part: the message getters (pass_up_m1_m2 / pass_down_ExxT of Gaussian, Addition, Multiplication, hstack) of the mirror
nodes at the random initial state of the shipped graph -- same seed => same state => same messages (node.py:95-129, 182-276;
nodes_todo.py:43-62; gaussian.py:179-183).  No kernel runs: before the first update() the nodes hold local numpy state.
GPU part (marked): src/tests.py:176-202 simple_PCA -- q = 1, W a single Gaussian column without an hstack -- through the
compiled plate, iteration by iteration."""
import numpy as np
import pytest

from helpers import load_golden, tensor_rel


def test_pyvb_alias_package_is_the_dropin():
    import pyvb
    import pyvb_b200
    from pyvb import nodes, Network                  # examples/PCA_missing_data.py:7, verbatim
    import pyvb.nodes as pn
    assert nodes is pyvb_b200.nodes and pn is nodes and Network is pyvb_b200.Network
    for name in ("Node", "Addition", "Multiplication", "Constant", "Gaussian", "DiagonalGaussian", "hstack", "Transpose",
                 "Gamma", "DiagonalGamma", "Wishart", "ConjugacyError"):
        assert hasattr(nodes, name), name


def test_messages_match_literal_reference():
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    from gen_golden_api import messages
    from pyvb import nodes
    g = load_golden("messages.npz")
    got = messages(nodes, seed=int(g["seed"]))       # the same builder, run against the mirror nodes
    assert set(got) == set(g)
    for k in sorted(g):
        assert got[k].shape == g[k].shape, k
        if k == "X":
            assert np.array_equal(got[k], g[k], equal_nan=True)
        elif np.max(np.abs(g[k])) > 0:
            assert tensor_rel(got[k], g[k]) < 1e-13, (k, tensor_rel(got[k], g[k]))
        else:
            assert np.all(got[k] == 0), k


@pytest.mark.gpu
def test_simple_pca_q1_without_hstack_matches_literal_reference():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    from gen_golden_api import build_simple_pca
    from pyvb import nodes
    g = load_golden("simple_pca.npz")
    np.random.seed(int(g["seed"]))
    N, d = g["X"].shape
    [np.random.randn(N, 1), np.random.randn(d, 1), np.random.randn(d, 1), np.random.randn(N, d)]   # the data draws
    noise, W, Mu, Zs, Xs = build_simple_pca(nodes, g["X"])
    assert tensor_rel(W.qmu[:, 0], g["init_W"]) == 0.0                  # the reference's random initialisation
    for it in range(int(g["niters"])):
        W.update()
        [e.update() for e in Zs]
        Mu.update()
        noise.update()
        p = "it%d_" % it
        assert tensor_rel(W.qmu[:, 0], g[p + "W"]) < 1e-9 and tensor_rel(W.qcov, g[p + "Wcov"]) < 1e-9
        assert tensor_rel(Mu.qmu[:, 0], g[p + "mu"]) < 1e-9 and tensor_rel(Mu.qcov, g[p + "mucov"]) < 1e-9
        assert tensor_rel(np.array([z.qmu[0, 0] for z in Zs]), g[p + "Z"]) < 1e-9
        assert tensor_rel(np.array([z.qcov[0, 0] for z in Zs]), g[p + "Zvar"]) < 1e-9
        assert abs(noise.qb - float(g[p + "qb"])) <= 1e-9 * float(g[p + "qb"])
    assert abs(noise.qa / noise.qb - float(g["it9_qa"]) / float(g["it9_qb"])) < 1e-9 * noise.qa / noise.qb


# ---------------------------------------------------------------------------------------------------------------------
# The LDS scripts through the node API (examples/Linear_Dynamic_System.py:47-76, LDS_knowns_in_A.py:72-74): the graph is
# compiled to the batched smoother, per-node update() calls are recorded and whole iterations become one launch.
def _lds_script(nodes, Y, q, known=None):
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    from gen_golden_lds import build

    class P(object):
        pass
    p = P()
    p.nodes = nodes
    m = build(p, Y, q)
    if known is not None:
        for i, a in enumerate(m["As"]):
            if not np.all(np.isnan(known[:, i])):
                a.observe(known[:, i].reshape(q, 1))
    return m


def _lds_iteration(m):
    Xs = m["Xs"]
    [x.update() for x in Xs]
    Xs.reverse()
    [x.update() for x in Xs]
    Xs.reverse()
    [a.update() for a in m["As"]]
    [c.update() for c in m["Cs"]]
    m["Q"].update()
    m["R"].update()


class _FakeLDS(object):
    calls = []

    def __init__(self, Y, q, **kw):
        self.Y, self.q, self.kw = np.asarray(Y), q, kw
        _FakeLDS.calls.append(("init", self.Y.shape, q, kw.get("A_known")))

    def set_state(self, st):
        self.st = {k: np.array(v) for k, v in st.items()}
        _FakeLDS.calls.append(("set_state", sorted(st)))

    def iterate(self, k):
        _FakeLDS.calls.append(("iterate", k))

    def check(self):
        pass

    def get_state(self):
        B, T, d = self.Y.shape
        out = dict(self.st)
        out["Xcov"] = np.zeros((B, T, self.q, self.q))
        out["Qa"], out["Ra"] = np.ones((B, self.q)), np.ones((B, d))
        return out


def test_lds_graph_compiles_to_one_launch_per_read(monkeypatch):
    from pyvb import nodes
    from pyvb_b200.lds_plate import LDSPlate
    monkeypatch.setattr(LDSPlate, "ENGINE", _FakeLDS)
    _FakeLDS.calls = []
    g = load_golden("lds_known.npz")
    np.random.seed(3)
    m = _lds_script(nodes, g["Y"], int(g["q"]), known=g["A_known"])
    for _ in range(3):
        _lds_iteration(m)
    assert [c[0] for c in _FakeLDS.calls] == ["init", "set_state"]     # compiled at the first update(), nothing launched yet
    A = np.hstack([a.qmu for a in m["As"]])                            # the first read runs the three recorded iterations
    kinds = [c[0] for c in _FakeLDS.calls]
    assert kinds == ["init", "set_state", "iterate"] and _FakeLDS.calls[-1] == ("iterate", 3)
    init = _FakeLDS.calls[0]
    assert init[1] == (1,) + g["Y"].shape and np.array_equal(init[3], g["A_known"], equal_nan=True)
    assert tensor_rel(A, g["init_A"]) == 0.0                           # (the fake engine returns the injected initial state)
    # reading again launches nothing; a partial sweep is refused
    _ = m["Q"].qb
    assert [c[0] for c in _FakeLDS.calls].count("iterate") == 1
    m["Xs"][0].update()
    m["Q"].update()
    with pytest.raises(NotImplementedError):
        _ = m["Xs"][0].qmu


def test_lds_plate_is_bound_lazily_and_keeps_the_reference_rng_order():
    """the mirror nodes draw their random initial state in the reference's order (incl. DiagonalGamma's single scalar)"""
    from pyvb import nodes
    g = load_golden("lds_a.npz")
    np.random.seed(0)
    m = _lds_script(nodes, g["Y"], int(g["q"]))
    assert tensor_rel(np.hstack([a.qmu for a in m["As"]]), g["init_A"]) == 0.0
    assert tensor_rel(np.hstack([c.qmu for c in m["Cs"]]), g["init_C"]) == 0.0
    assert float(m["Q"].qb) == float(g["init_Qb"][0]) and float(m["R"].qb) == float(g["init_Rb"][0])
    assert tensor_rel(np.stack([x.qmu[:, 0] for x in m["Xs"]]), g["init_X"]) == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("name,seed", [("lds_a.npz", 0), ("lds_b.npz", 1), ("lds_known.npz", 3), ("lds_known_b.npz", 4)])
def test_lds_scripts_through_the_node_api_match_literal_reference(name, seed):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from pyvb import nodes
    g = load_golden(name)
    q = int(g["q"])
    np.random.seed(seed)
    m = _lds_script(nodes, g["Y"], q, known=g["A_known"] if "A_known" in g else None)
    for it in range(int(g["niters"])):
        _lds_iteration(m)
        if it % 2 == 0:
            continue                                                    # two iterations in one launch every other time
        p = "it%d_" % it
        assert tensor_rel(np.hstack([a.qmu for a in m["As"]]), g[p + "A"]) < 1e-9, (name, it)
        assert tensor_rel(np.stack([np.diag(a.qcov) for a in m["As"]], 1), g[p + "Avar"]) < 1e-9 or np.max(g[p + "Avar"]) == 0
        assert tensor_rel(np.hstack([c.qmu for c in m["Cs"]]), g[p + "C"]) < 1e-9
        assert tensor_rel(np.stack([x.qmu[:, 0] for x in m["Xs"]]), g[p + "X"]) < 1e-9
        assert tensor_rel(np.stack([x.qcov for x in m["Xs"]]), g[p + "Xcov"]) < 1e-9
        assert tensor_rel(m["Q"].qb, g[p + "Qb"]) < 1e-9 and tensor_rel(m["R"].qb, g[p + "Rb"]) < 1e-9
        assert tensor_rel(np.diag(m["Q"].pass_down_Ex()), g[p + "Qa"] / g[p + "Qb"]) < 1e-9
    b = m["Q"]._binding
    assert b.iterations == int(g["niters"]) and b.launches == int(g["niters"]) // 2


@pytest.mark.gpu
def test_simple_regression_matches_literal_reference():
    """src/tests.py:100-128: constant scalar regressors on the left of the product, manual order A, B, noise."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    from gen_golden_api import build_regression
    from pyvb import nodes
    g = load_golden("simple_regression.npz")
    np.random.seed(int(g["seed"]))
    np.random.randn(g["x"].shape[0], 1)                                 # the data draw
    A, B, noise, Ys = build_regression(nodes, g["x"], g["y"])
    got = np.array([A.qmu[0, 0], A.qcov[0, 0], B.qmu[0, 0], B.qcov[0, 0], noise.qb])
    assert np.array_equal(got, g["init"])                                # the reference's random initialisation
    for it in range(int(g["niters"])):
        A.update()
        B.update()
        noise.update()
        got = np.array([A.qmu[0, 0], A.qcov[0, 0], B.qmu[0, 0], B.qcov[0, 0], noise.qb, noise.pass_down_Ex()[0, 0]])
        assert np.max(np.abs(got - g["it%d" % it]) / np.abs(g["it%d" % it])) < 1e-9, (it, got, g["it%d" % it])
