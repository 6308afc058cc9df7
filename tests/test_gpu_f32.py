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
    ncp, zoff = int(lib.pyvb_f32_pitch(q)), int(lib.pyvb_f32_zoff(q))
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
    assert tensor_rel(out[:, :P], qprec) < 2e-6
    assert tensor_rel(out[:, zoff:zoff + q], eta) < 2e-5
    assert np.all(out[:, P:zoff] == 0) and np.all(out[:, zoff + q:] == 0)
