"""FP32 variant (tcgen05 / TMEM / TMA) through the C-ABI against float64 numpy.

K1-f32 computes, per row, [qprec | eta] = [P0 | h0] + tau * (mask @ G + (mask * (x - mu)) @ W) -- the contraction of
Multiplication.pass_up_m1_m2 (/root/reference/src/pyvb/nodes/node.py:203-227) with a masked precision -- from bf16
splits: the mask product is FP32-exact (3-way split of G), x carries a 16-bit mantissa (2-way split).
Stated tolerance (tensor-wise relative, vs float64): 2e-6 on qprec, 2e-5 on eta."""
import numpy as np
import pytest

from helpers import tensor_rel
from oracle.plate_oracle import synth_pca

pytestmark = pytest.mark.gpu


def _ref(X, Wbar, Wvar, mu, tau, q):
    O = ~np.isnan(X)
    X0 = np.where(O, X, 0.0)
    ii, jj = np.tril_indices(q)
    G = Wbar[:, ii] * Wbar[:, jj] + np.where(ii == jj, Wvar[:, ii], 0.0)           # D x P
    qprec = (ii == jj).astype(float)[None, :] + tau * (O.astype(float) @ G)
    eta = tau * ((X0 - O * mu[None, :]) @ Wbar)
    return qprec, eta


@pytest.mark.parametrize("shape", [(128, 32, 16), (1000, 256, 16), (300, 64, 32), (515, 1024, 32), (260, 96, 64)])
def test_k1_f32_matches_float64(shape):
    import torch
    from pyvb_b200 import _cabi
    lib = _cabi.lib()
    N, D, q = shape
    assert lib.pyvb_f32_supported(D, q)
    dev = torch.device("cuda", 0)
    X = synth_pca(N, D, q, 0.3, seed=N + D)
    if N > 10:
        X[2, :] = np.nan
        X[5, :] = 1.0
    rng = np.random.RandomState(q)
    Wbar, Wvar, mu, tau = rng.randn(D, q), rng.rand(D, q) + 0.1, rng.randn(D) * 0.3, 7.5
    ncp, zoff, poff = int(lib.pyvb_f32_pitch(q)), int(lib.pyvb_f32_zoff(q)), int(lib.pyvb_f32_poff(q))
    P = q * (q + 1) // 2
    st = torch.cuda.current_stream(dev).cuda_stream
    Xd = torch.as_tensor(X, device=dev)
    planes = torch.zeros(3, N, D, dtype=torch.bfloat16, device=dev)
    _cabi.check(lib.pyvb_prepare_x_f32(N, D, Xd.data_ptr(), D, planes.data_ptr(), st), "prepare_x")
    O = ~torch.isnan(Xd)
    assert torch.equal(planes[0].float(), O.float())
    x0 = torch.where(O, Xd, torch.zeros((), dtype=Xd.dtype, device=dev))
    assert float((planes[1].double() + planes[2].double() - x0).abs().max()) <= 2.0 ** -15 * float(x0.abs().max())
    W_t, V_t, mu_t = (torch.as_tensor(a, device=dev) for a in (Wbar, Wvar, mu))
    GT = torch.zeros(3, ncp, D, dtype=torch.bfloat16, device=dev)
    WT = torch.zeros(3, q, D, dtype=torch.bfloat16, device=dev)
    _cabi.check(lib.pyvb_pack_gw_f32(D, q, W_t.data_ptr(), V_t.data_ptr(), mu_t.data_ptr(), GT.data_ptr(), WT.data_ptr(), st),
                "pack_gw_f32")
    assert float((WT.double().sum(0).t() - W_t).abs().max()) <= 2.0 ** -22 * float(W_t.abs().max())
    P0 = torch.eye(q, dtype=torch.float64, device=dev)
    h0 = torch.zeros(q, dtype=torch.float64, device=dev)
    gl = torch.zeros(144, dtype=torch.float64, device=dev)
    gl[2] = tau
    MZ = torch.full((N, ncp), float("nan"), dtype=torch.float32, device=dev)
    _cabi.check(lib.pyvb_zstep_k1_f32(N, D, q, planes.data_ptr(), GT.data_ptr(), WT.data_ptr(), P0.data_ptr(), h0.data_ptr(),
                                      gl.data_ptr(), MZ.data_ptr(), st), "zstep_k1_f32")
    torch.cuda.synchronize()
    out = MZ.cpu().numpy().astype(np.float64)
    qprec, eta = _ref(X, Wbar, Wvar, mu, tau, q)
    assert np.all(np.isfinite(out))
    assert tensor_rel(out[:, poff:poff + P], qprec) < 2e-6
    assert tensor_rel(out[:, zoff:zoff + q], eta) < 2e-5
    assert np.all(out[:, poff + P:] == 0)


F32_TOL_STEP = 5e-5       # stated tolerance of the FP32 variant: tensor-wise relative error of ONE sweep started from
                          # the same state as the float64 oracle (bf16 x 3 contraction with FP32 tensor-core accumulation)
F32_TOL_TRAJ = 5e-3       # ... and of 5 free-running sweeps from a random initialisation (the VB map amplifies the
                          # one-step error transiently)
F32_TOL_QB = 5e-3         # qb (hence tau) is a difference of sums 100-500x its size: the one-step error is amplified
F32_TOL_ELBO = 1e-3       # relative error of the bound after one sweep from the same state (dominated by tau * resid)


@pytest.mark.parametrize("shape", [(3000, 256, 16), (1200, 64, 32), (130, 32, 16), (4100, 1024, 32)])
def test_f32_engine_sweeps_track_the_oracle(shape):
    """Whole sweeps of the FP32 variant (tcgen05 contraction + FP64-internal batched solve on FP32 rows + tcgen05
    statistics) against the float64 oracle: one-step error from a shared state, and the free-running trajectory."""
    from pyvb_b200 import PlateEngine
    from oracle.plate_oracle import PlateOracle
    N, D, q = shape
    X = synth_pca(N, D, q, 0.25, seed=N)
    X[1, :] = np.nan
    rng = np.random.RandomState(11)
    init = {"Wbar": rng.randn(D, q), "Wvar": np.ones((D, q)), "mu": np.zeros(D), "muvar": np.ones(D),
            "Zbar": rng.randn(N, q), "Sig": np.tile(np.eye(q), (N, 1, 1)), "qb": 0.5}
    free = PlateOracle(X, q, mode="B")
    free.load_state(init)
    e = PlateEngine(X, q, mode="B", device="cuda:0", precision="f32")
    e.set_state(init)
    keys = ("Wbar", "Wvar", "mu", "Zbar", "Sig")
    worst = {}
    for it in range(5):
        st0 = e.get_state()                              # the state the FP32 engine actually holds
        step = PlateOracle(X, q, mode="B")
        step.load_state({k: st0[k] for k in ("Wbar", "Wvar", "mu", "muvar", "Zbar", "Sig", "qb")})
        ref_free, ref, got = free.iterate(), step.iterate(), e.iterate()
        st = e.get_state()
        for k in keys:
            err = tensor_rel(st[k], getattr(step, k))
            worst[k] = max(worst.get(k, 0.0), err)
            assert err < F32_TOL_STEP, (shape, it, k, err)
        worst["qb"] = max(worst.get("qb", 0.0), abs(st["qb"] - step.qb) / abs(step.qb))
        assert abs(st["qb"] - step.qb) <= F32_TOL_QB * abs(step.qb)
        if np.isfinite(ref):
            worst["elbo"] = max(worst.get("elbo", 0.0), abs(got - ref) / abs(ref))
            assert abs(got - ref) <= F32_TOL_ELBO * abs(ref), (shape, it, got, ref)
    for k in keys:
        assert tensor_rel(st[k], getattr(free, k)) < F32_TOL_TRAJ, (shape, k)
    e.check()
    print("f32 worst one-step errors", shape, {k: "%.1e" % v for k, v in worst.items()})
