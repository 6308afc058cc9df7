"""GPU parity of the INT8 tensor-core mask contraction (kernels_i8.cu, pyvb_zstep_i8_f64): the same FP64 results as
the DMMA path and the oracle, through the C-ABI.  The digit split is exact by construction (integer GEMMs), so the
tolerance is the one of the other FP64 kernels: 1e-11 kernel against kernel, 1e-9 against the oracle per sweep."""
import numpy as np
import pytest
import torch

from helpers import tensor_rel
from oracle.plate_oracle import PlateOracle, synth_pca
from test_gpu_parity import rand_init, _cmp_state, TOL

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from pyvb_b200 import PlateEngine
    return PlateEngine


I8_SHAPES = [(5000, 256, 16), (64, 64, 16), (1, 64, 16), (777, 128, 32), (2100, 192, 32), (333, 64, 64), (1500, 128, 64),
             (19000, 1024, 32)]


@pytest.mark.parametrize("shape", I8_SHAPES)
def test_i8_zstep_matches_dmma(eng, shape):
    """kernel-level: the MZ rows, Sigma and log-dets of one Z step, INT8 mask contraction against the all-DMMA path."""
    N, D, q = shape
    X = synth_pca(N, D, q, 0.3, seed=N + q)
    if N > 10:
        X[2, :] = np.nan            # an all-missing row
        X[5, :] = 1.0               # a fully observed row
    init = rand_init(N, D, q, seed=7)
    init["Wbar"][:, 0] *= 1e3       # columns of G on very different scales (per-column fixed-point scale)
    init["Wbar"][:, 1] *= 1e-4
    ed, ei = eng(X, q, mode="B", algo="dmma"), eng(X, q, mode="B", algo="i8")
    assert ei.use_i8 and not ed.use_i8
    for e in (ed, ei):
        e.set_state(init)
        e._ensure_stats()           # fills the cache of the X-only sums (the INT8 statistics need it)
        e.update_Z()
        e._ensure_stats()
    assert getattr(ei, "i8_stats_calls", 0) == 1 and ei.use_i8_stats
    sd, si = ed.get_state(), ei.get_state()
    for k in ("Zbar", "Sig"):
        assert tensor_rel(si[k], sd[k]) < 1e-11, (shape, k)
    assert tensor_rel(ei.M2.contiguous().cpu().numpy(), ed.M2.contiguous().cpu().numpy()) < 1e-11
    ld, li = ed.logdet.cpu().numpy(), ei.logdet.cpu().numpy()
    assert np.max(np.abs(ld - li)) < 1e-11 * max(1.0, np.max(np.abs(ld)))
    vd, vi = ed.L.views(ed.stats.cpu().numpy()), ei.L.views(ei.stats.cpu().numpy())
    for k in ("T1", "Bst", "Ast", "cnt", "colx", "S", "zsum"):
        assert tensor_rel(vi[k], vd[k]) < 1e-11, (shape, k, tensor_rel(vi[k], vd[k]))
    ei.check()


@pytest.mark.parametrize("shape", [(600, 64, 16), (300, 128, 32)])
def test_i8_qprec_is_exact_to_rounding(eng, shape):
    """The qprec columns alone (k1_only), against a float128-free exact check: every G entry is rounded to
    scale * 2^-55, the integer sums are exact, so |qprec_i8 - qprec_exact| <= tau * n_obs * scale_c * 2^-55 + 1 ulp."""
    from pyvb_b200 import _cabi
    N, D, q = shape
    X = synth_pca(N, D, q, 0.4, seed=3)
    init = rand_init(N, D, q, seed=5)
    e = eng(X, q, mode="B", algo="i8")
    e.set_state(init)
    e._ensure_gw()
    lib = e.lib
    _cabi.check(lib.pyvb_prepare_mask_i8(N, D, e.X.data_ptr(), D, e.mask8.data_ptr(), e._stream()), "mask")
    rc = lib.pyvb_zstep_i8_f64(N, D, q, e.X.data_ptr(), D, e.mask8.data_ptr(), e.Wbar.data_ptr(), e.Wvar.data_ptr(),
                               e.Gw.data_ptr(), e.ldg, e.P0.data_ptr(), e.h0.data_ptr(), e.gl.data_ptr(),
                               e.MZ.data_ptr(), e.ldmz, e.GI.data_ptr(), e.gscale.data_ptr(), 0, e.logdet.data_ptr(), 0,
                               1, e._stream())
    _cabi.check(rc, "zstep_i8 k1")
    torch.cuda.synchronize()
    P = q * (q + 1) // 2
    got = e.MZ[:, :P].cpu().numpy()
    eta = e.MZ[:, e.zoff:e.zoff + q].cpu().numpy()
    O = (~np.isnan(X))
    ii, jj = np.tril_indices(q)
    W, Wv = init["Wbar"], init["Wvar"]
    G = W[:, ii] * W[:, jj] + np.where(ii == jj, Wv[:, ii], 0.0)                  # D x P
    tau = e.get_state()["tau"]
    ref = np.eye(q)[ii, jj][None, :] + tau * (O.astype(np.longdouble) @ G.astype(np.longdouble)).astype(np.float64)
    from oracle.i8_oracle import column_scales, mask_contract_i8
    scale = column_scales(G)                                   # the power of two above the column maximum
    bound = tau * O.sum(1)[:, None] * scale[None, :] * 2.0 ** -55 + 4 * np.finfo(np.float64).eps * np.abs(ref)
    assert np.all(np.abs(got - ref) <= bound), float(np.max(np.abs(got - ref) / bound))
    # the numpy restatement of the kernel steps (oracle/i8_oracle.py) agrees to the last rounding (FMA vs multiply + add)
    rest = mask_contract_i8(O, G, tau=tau, add=np.eye(q)[ii, jj][None, :])
    assert np.all(np.abs(got - rest) <= 2 * np.finfo(np.float64).eps * np.abs(rest))
    X0 = np.where(O, X - init["mu"][None, :], 0.0)
    assert tensor_rel(eta, tau * (X0 @ W)) < 1e-12
    m = e.mask8.cpu().numpy().reshape(-1, D // 64, 128, 64).transpose(0, 2, 1, 3).reshape(-1, D)   # tiles -> rows
    assert np.array_equal(m[:N].astype(bool), O) and not m[N:].any()


@pytest.mark.parametrize("shape", [(3000, 256, 16), (1200, 64, 32), (700, 128, 64)])
def test_i8_iterations_match_oracle(eng, shape):
    N, D, q = shape
    X = synth_pca(N, D, q, 0.25, seed=N)
    init = rand_init(N, D, q, seed=11)
    o = PlateOracle(X, q, mode="B")
    o.load_state(init)
    e = eng(X, q, mode="B", algo="i8")
    e.set_state(init)
    for it in range(5):
        ref, got = o.iterate(), e.iterate()
        st = e.get_state()
        _cmp_state(st, o.state(), ("Wbar", "Wvar", "mu", "muvar", "Zbar", "Sig"), (shape, it))
        assert abs(st["qb"] - o.qb) <= TOL * abs(o.qb)
        assert abs(got - ref) <= TOL * abs(ref), (shape, it, got, ref)
    assert getattr(e, "i8_stats_calls", 0) >= 4
    e.check()


@pytest.mark.parametrize("shape", [(129, 64, 16), (257, 128, 16), (385, 1280, 16), (130, 1024, 32), (300, 1024, 64)])
def test_i8_edge_shapes_one_sweep(eng, shape):
    """Odd numbers of 128-row blocks (a cluster's second CTA runs past the end), the largest supported D, tiny N."""
    N, D, q = shape
    X = synth_pca(N, D, q, 0.3, seed=D + N)
    init = rand_init(N, D, q, seed=3)
    ed, ei = eng(X, q, mode="B", algo="dmma"), eng(X, q, mode="B", algo="i8")
    for e in (ed, ei):
        e.set_state(init)
        e._ensure_stats()
    vals = [(ed.iterate(), ei.iterate()) for _ in range(2)]
    for a, b in vals:
        assert abs(a - b) <= 1e-10 * abs(a), (shape, vals)
    sd, si = ed.get_state(), ei.get_state()
    for k in ("Wbar", "mu", "Zbar", "Sig"):
        assert tensor_rel(si[k], sd[k]) < 1e-10, (shape, k)
    ei.check()


def test_i8_auto_falls_back_where_the_mask_block_does_not_fit(eng):
    X = synth_pca(300, 1280, 64, 0.2, seed=2)          # q = 64 at D = 1280: constants + mask block exceed shared memory
    e = eng(X, 64, mode="B")
    assert not e.use_i8
    with pytest.raises(ValueError):
        eng(X, 64, mode="B", algo="i8")


def test_i8_rejects_unsupported_shape(eng):
    X = synth_pca(40, 48, 16, 0.1, seed=1)
    with pytest.raises(ValueError):
        eng(X, 16, mode="B", algo="i8")


# ---------------------------------------------------------------------------------------------------------------------
# Dynamic range: the fixed point of the INT8 kernels is relative to the COLUMN maximum (of G over the data dimensions, of
# the MZ columns over the rows).  An un-normalised feature (one ROW of W on another scale) or outlier data rows leave the
# rows / dimensions that do not see them with an absolute error of scale * 2^-55 in a result of ordinary size.  The default
# path (algo="auto") must stay within 1e-9 PER ROW there: the guard sends such steps to the FP64 tensor cores.
def _per_row_rel(a, b):
    a = np.asarray(a).reshape(a.shape[0], -1)
    b = np.asarray(b).reshape(b.shape[0], -1)
    return np.max(np.abs(a - b), axis=1) / np.maximum(np.max(np.abs(b), axis=1), 1e-300)


@pytest.mark.parametrize("s", [1.0, 30.0, 1e4, 1e6])
@pytest.mark.parametrize("shape", [(2000, 128, 16), (900, 192, 32)])
def test_auto_path_is_row_accurate_with_an_unnormalised_feature(eng, shape, s):
    """Rows that do NOT observe the scaled feature are the hazard (their qprec is of ordinary size, the fixed-point unit is
    s^2 times too coarse): 1e-9 per row against the oracle.  Rows that do observe it have cond(qprec) ~ s^2, so any two
    FP64 evaluations differ by ~cond eps there: the default path must be as accurate as the all-DMMA path."""
    N, D, q = shape
    X = synth_pca(N, D, q, 0.3, seed=21)
    X[:, 9] *= s                                   # feature 9 lives on another scale ...
    init = rand_init(N, D, q, seed=4)
    init["Wbar"][9] *= s                           # ... and so does its row of W
    hazard = np.isnan(X[:, 9])
    o = PlateOracle(X, q, mode="B")
    o.load_state(init)
    e, ed = eng(X, q, mode="B", algo="auto"), eng(X, q, mode="B", algo="dmma")
    assert e.use_i8 and e.use_i8_stats
    for g in (e, ed):
        g.set_state(init)
        g._ensure_stats()
        g.update_Z()
    o.update_Z()
    st, sd = e.get_state(), ed.get_state()
    for k in ("Sig", "Zbar"):
        ea, edm = _per_row_rel(st[k], getattr(o, k)), _per_row_rel(sd[k], getattr(o, k))
        assert ea[hazard].max() < 1e-9, (shape, s, k, ea[hazard].max())
        assert np.all((ea < 1e-9) | (ea <= 8 * edm)), (shape, s, k, float(np.max(ea / np.maximum(edm, 1e-300))))
    zf, _ = e.i8_fallbacks()
    assert (zf >= 1) == (s >= 1e4), (s, zf)        # the guard fires exactly where the hazard is
    for it in range(3):                            # and whole sweeps stay on the oracle (as well as the DMMA path does)
        ref, got, gd = o.iterate(), e.iterate(), ed.iterate()
        assert abs(got - ref) <= max(TOL * abs(ref), 8 * abs(gd - ref)), (shape, s, it, got, gd, ref)
    st, sd = e.get_state(), ed.get_state()
    for k in ("Wbar", "Sig"):
        ea, edm = _per_row_rel(st[k], getattr(o, k)), _per_row_rel(sd[k], getattr(o, k))
        assert np.all((ea < 1e-9) | (ea <= 8 * edm)), (shape, s, k)
    e.check()


@pytest.mark.parametrize("s", [1.0, 1e4])
def test_auto_path_is_accurate_per_dimension_with_outlier_rows(eng, s):
    N, D, q = 3000, 128, 16
    X = synth_pca(N, D, q, 0.3, seed=8)
    X[:6] *= s                                     # six outlier rows ...
    X[:6, :20] = np.nan                            # ... which the first 20 data dimensions never see
    init = rand_init(N, D, q, seed=9)
    o = PlateOracle(X, q, mode="B")
    o.load_state(init)
    e = eng(X, q, mode="B", algo="auto")
    e.set_state(init)
    e._ensure_stats()
    o.update_Z(); e.update_Z()
    e._stats_fresh = False
    e._ensure_stats()
    from helpers import numpy_stats
    ref = e.L.views(numpy_stats(X, o.Zbar, o.Sig, q))
    got = e.L.views(e.stats.cpu().numpy())
    for k in ("T1", "Bst", "Ast"):
        assert _per_row_rel(got[k], ref[k]).max() < 1e-9, (k, s, _per_row_rel(got[k], ref[k]).max())
    _, sf = e.i8_fallbacks()
    assert (sf >= 1) == (s >= 1e4), (s, sf)
    for it in range(3):
        r, g = o.iterate(), e.iterate()
        assert abs(g - r) <= TOL * abs(r), (s, it, g, r)
    assert _per_row_rel(e.get_state()["Wbar"], o.Wbar).max() < 1e-9
    e.check()


def test_forced_fallback_reproduces_the_dmma_path(eng):
    """PYVB_I8_GUARD=force (read once per process, hence the subprocess) makes every guard fire: the conditional launches must
    then leave exactly what the all-DMMA path computes (up to the order of the chunk sums of the statistics)."""
    import subprocess, sys, os, textwrap
    code = textwrap.dedent("""
        import numpy as np, torch, sys
        sys.path.insert(0, %r)
        from pyvb_b200 import PlateEngine
        from oracle.plate_oracle import synth_pca
        X = synth_pca(1500, 128, 32, 0.3, seed=1)
        ed, ei = PlateEngine(X, 32, mode="B", algo="dmma"), PlateEngine(X, 32, mode="B", algo="auto")
        for e in (ed, ei):
            e.init_random(seed=3)
            for _ in range(3):
                e.iterate()
        assert ei.i8_fallbacks() == (3, 3), ei.i8_fallbacks()
        rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
        assert rel(ei.MZ, ed.MZ) < 1e-11 and rel(ei.Wbar, ed.Wbar) < 1e-11 and rel(ei.trace[:3], ed.trace[:3]) < 1e-11
        e2 = PlateEngine(X, 32, mode="B", algo="auto")          # one Z step from the same state: the rows are bit-identical
        e2.set_state(ed.get_state()); ed.set_state(ed.get_state())
        e2._ensure_stats(); ed._ensure_stats()
        e2.update_Z(); ed.update_Z()
        assert torch.equal(e2.MZ, ed.MZ) and torch.equal(e2.logdet, ed.logdet)
        print("forced ok")
    """ % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    env = dict(os.environ, PYVB_I8_GUARD="force")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "forced ok" in r.stdout, r.stdout + r.stderr
