"""TEST INFRASTRUCTURE (never imported by pyvb_b200/): numpy restatement of the arithmetic of the Gauss-Jordan batched solve
(pyvb_b200/csrc/kernels_k2g.cu), the kernel that replaces cho_factor / cho_solve(., I) / dot(qcov, .) / q_ln_det of
Gaussian.update (/root/reference/src/pyvb/nodes/gaussian.py:117-123).

Row i of every matrix is one lane's registers; step k publishes  b_i = a[i][k] * sinv_i  (sinv_i = 1 before row i's pivot,
-1 / d_i after it), every row adds  t_i * b_j  (t_i = -a[i][k] / d_k, t = 0 for the pivot row, which gets a 1 in column k) and the
rows are scaled by 1 / d_i only at the end.  Same operation order as the kernel (one rounding per fused multiply-add is NOT
modelled: numpy rounds the product first; the difference is one ulp per update)."""
import numpy as np


def gj_solve(A, eta):
    """A [N, q, q] SPD, eta [N, q] -> Sigma = A^-1, zbar = Sigma eta, ln prod diag chol(A) (all batched)."""
    A = np.asarray(A, dtype=np.float64)
    N, q, _ = A.shape
    a = A.copy()
    sinv = np.ones((N, q))
    for k in range(q):
        c = a[:, :, k].copy()
        b = c * sinv                                   # the published column
        rc = 1.0 / b[:, k]
        t = -c * rc[:, None]
        t[:, k] = 0.0
        upd = t[:, :, None] * b[:, None, :]
        upd[:, :, k] = 0.0
        a += upd
        a[:, :, k] = t
        a[:, k, k] = 1.0
        sinv[:, k] = -rc
    Sigma = a * (-sinv)[:, :, None]
    with np.errstate(invalid="ignore", divide="ignore"):
        logdet = -0.5 * np.sum(np.log(-sinv), axis=1)
    z = np.einsum("nij,nj->ni", Sigma, eta)
    return Sigma, z, logdet
