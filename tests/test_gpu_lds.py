"""LDS VB smoother (BASELINE config 5) on the GPU through the C-ABI (pyvb_lds_iterate_f64):
* against the LITERAL reference (fixtures of oracle/gen_golden_lds.py: examples/Linear_Dynamic_System.py:47-76 run
  with a seed), iteration by iteration, 1e-9;
* against the batched numpy restatement on many independent sequences, incl. several iterations in one launch."""
import numpy as np
import pytest

from helpers import load_golden, tensor_rel
from oracle.lds_oracle import LDSOracle, synth_lds

pytestmark = pytest.mark.gpu
TOL = 1e-9
KEYS = ("A", "Avar", "C", "Cvar", "Qa", "Qb", "Ra", "Rb", "X", "Xcov")


@pytest.mark.parametrize("name", ["lds_a.npz", "lds_b.npz", "lds_c.npz"])
def test_lds_matches_literal_reference(name):
    from pyvb_b200 import LDSEngine
    g = load_golden(name)
    e = LDSEngine(g["Y"], int(g["q"]), device="cuda:0")
    e.set_state({k: g["init_" + k] for k in LDSEngine.KEYS})
    for it in range(int(g["niters"])):
        e.iterate()
        st = e.get_state()
        for k in KEYS:
            assert tensor_rel(st[k][0], g["it%d_%s" % (it, k)]) < TOL, (name, it, k)
    e.check()


@pytest.mark.parametrize("shape", [(300, 40, 8, 5), (1000, 17, 3, 2), (5, 200, 8, 8), (64, 9, 1, 1), (7, 3, 2, 2), (3, 67, 8, 5),
                                   (2, 131, 4, 3)])
def test_lds_batch_matches_oracle(shape):
    from pyvb_b200 import LDSEngine
    B, T, q, d = shape
    Y = synth_lds(B, T, q, d, seed=B + T)
    rng = np.random.RandomState(3)
    init = {"A": rng.randn(B, q, q) * 0.3, "Avar": rng.rand(B, q, q) + 0.5, "C": rng.randn(B, d, q),
            "Cvar": rng.rand(B, d, q) + 0.5, "Qb": rng.rand(B, q) + 0.2, "Rb": rng.rand(B, d) + 0.2,
            "X": rng.randn(B, T, q)}
    o = LDSOracle(Y, q)
    o.load_state(init)
    e = LDSEngine(Y, q, device="cuda:0")
    e.set_state(init)
    e2 = LDSEngine(Y, q, device="cuda:0")
    e2.set_state(init)
    for it in range(4):
        o.iterate()
        e.iterate()
        st = e.get_state()
        for k in KEYS:
            assert tensor_rel(st[k], getattr(o, k)) < TOL, (shape, it, k)
    e2.iterate(4)                                   # the same four iterations inside one launch
    st2 = e2.get_state()
    for k in KEYS:
        assert np.array_equal(st2[k], st[k]), k
    e.check()


def test_lds_scan_and_step_by_step_sweeps_agree(monkeypatch):
    """The chunked-scan sweeps (default) against the step-by-step Gauss-Seidel sweeps of round 1 (PYVB_LDS=serial): the same
    arithmetic up to the rounding of the chunk starts."""
    from pyvb_b200 import LDSEngine
    B, T, q, d = 40, 200, 8, 5
    Y = synth_lds(B, T, q, d, seed=11)
    out = {}
    for mode in ("scan", "serial"):
        monkeypatch.setenv("PYVB_LDS", mode)
        e = LDSEngine(Y, q, device="cuda:0")
        e.init_random(seed=2)
        e.iterate(3)
        out[mode] = e.get_state()
        e.check()
    for k in KEYS:
        assert tensor_rel(out["scan"][k], out["serial"][k]) < 1e-12, k


def test_lds_empty_batch_is_a_noop():
    from pyvb_b200 import LDSEngine
    e = LDSEngine(np.zeros((0, 10, 3)), 2, device="cuda:0")
    e.iterate(2)
    e.check()
    assert e.get_state()["X"].shape == (0, 10, 2)


@pytest.mark.parametrize("name", ["lds_known.npz", "lds_known_b.npz"])
def test_lds_known_entries_of_A_match_literal_reference(name):
    """examples/LDS_knowns_in_A.py:72-74: columns of A observed with NaN = unknown."""
    from pyvb_b200 import LDSEngine
    g = load_golden(name)
    e = LDSEngine(g["Y"], int(g["q"]), device="cuda:0", A_known=g["A_known"])
    e.set_state({k: g["init_" + k] for k in LDSEngine.KEYS})
    for it in range(int(g["niters"])):
        e.iterate()
        st = e.get_state()
        for k in KEYS:
            assert tensor_rel(st[k][0], g["it%d_%s" % (it, k)]) < TOL, (name, it, k)
        kn = g["A_known"]
        assert np.array_equal(st["A"][0][~np.isnan(kn)], kn[~np.isnan(kn)]) and np.all(st["Avar"][0][~np.isnan(kn)] == 0)
    e.check()
