"""TEST INFRASTRUCTURE (never the product): numpy restatement of the fixed-point scheme behind the INT8 tensor-core kernels
(pyvb_b200/csrc/kernels_i8.cu), i.e. of how the reference's mask contraction

    m1 = tr(<w_i w_j^T> Lambda_n)  with  Lambda_n = tau diag(mask_n)        /root/reference/src/pyvb/nodes/node.py:213-227

is evaluated exactly with integer GEMMs:  G[d][c] = scale_c 2^-54 sum_t digit_t[d][c] 256^t  (seven balanced base-256 digits),
mask @ G = scale_c 2^-54 sum_t 256^t (mask @ digit_t).  Every step below mirrors one kernel step (pack_g_i8_kernel /
digitize_kernel, the tcgen05.mma kind::i8 products, combine7 in the epilogue); the tests pin it against exact integer
arithmetic (Python ints) and against a long-double product.
"""
import numpy as np

NPL = 7
TWO54 = float(2 ** 54)
BIAS = sum(128 << (8 * t) for t in range(NPL))          # 0x0080808080808080


def column_scales(G):
    """scale_c = the power of two above max_d |G[d][c]| (1 for an all-zero column), as pow2_above() in the kernels: the
    scaling by 2^54 / scale_c is then exact."""
    s = np.max(np.abs(G), axis=0)
    _, e = np.frexp(s)
    return np.where((s > 0) & (s < 1e300), np.ldexp(1.0, e), 1.0)


def to_fixed(G, scale):
    """round-to-nearest-even of G * (2^54 / scale_c) as int64 (|v| <= 2^54), the kernels' __double2ll_rn(g * inv)."""
    inv = TWO54 / scale
    return np.rint(G * inv[None, :]).astype(np.int64)


def balanced_digits(v):
    """Seven balanced base-256 digits (int8, [-128, 127]) of int64 v, |v| < 2^55: the bytes of v + BIAS with bit 7 flipped."""
    u = (v.astype(np.int64) + np.int64(BIAS)).astype(np.uint64)
    planes = []
    for t in range(NPL):
        b = ((u >> np.uint64(8 * t)) & np.uint64(0xFF)).astype(np.int64)
        planes.append((b - 128).astype(np.int8))
    return np.stack(planes)                                   # [7][...]


def combine7(acc):
    """The epilogue's recombination of the seven INT32 accumulators acc[7][...]: two exact 64-bit integer halves, each
    converted exactly to double, one rounding in hi * 2^32 + lo."""
    a = acc.astype(np.int64)
    lo = a[0] + (a[1] << 8) + (a[2] << 16) + (a[3] << 24)
    hi = a[4] + (a[5] << 8) + (a[6] << 16)
    assert np.all(np.abs(lo) < 2 ** 51) and np.all(np.abs(hi) < 2 ** 51)
    return hi.astype(np.float64) * 4294967296.0 + lo.astype(np.float64)     # the multiplication is exact: one rounding


def mask_contract_i8(mask, G, tau=1.0, add=None):
    """add + tau * mask @ G evaluated as the kernels do.  mask: [N][D] 0/1, G: [D][C] float64."""
    scale = column_scales(G)
    dig = balanced_digits(to_fixed(G, scale))                                  # [7][D][C]
    m = mask.astype(np.int32)
    acc = np.stack([m @ dig[t].astype(np.int32) for t in range(NPL)])          # exact INT32 GEMMs
    assert np.all(np.abs(acc.astype(np.int64)) < 2 ** 31)
    f = (tau * scale) * 5.5511151231257827e-17                                 # tau * scale_c * 2^-54
    out = f[None, :] * combine7(acc)
    return out if add is None else out + add


def exact_fixed_sum(mask, G):
    """sum_d mask[n][d] round(G[d][c] 2^54 / scale_c) in exact Python integer arithmetic ([N][C] list of ints) + the scales."""
    scale = column_scales(G)
    v = to_fixed(G, scale)
    N, C = mask.shape[0], G.shape[1]
    out = [[0] * C for _ in range(N)]
    vl = v.tolist()
    for n in range(N):
        idx = np.nonzero(mask[n])[0].tolist()
        for c in range(C):
            out[n][c] = sum(vl[d][c] for d in idx)
    return out, scale


# ---- the accuracy guard (pyvb_b200/csrc/cabi.cu: I8_TOL; kernels.h: I8Check; kernels_i8.cu: stats_i8_check_kernel) ----
I8_TOL = 2.0 ** -38


def zstep_guard_rows(qprec_diag_max, G, tau, D, tol=I8_TOL):
    """Rows of a Z step the batched solve flags: the fixed-point bound  tau * D * max_c scale_c * 2^-55  exceeds
    tol * max_i qprec_ii.  qprec_diag_max: [N] largest diagonal entry of every row's posterior precision."""
    thr = tau * (D * 2.0 ** -55 / tol) * float(np.max(column_scales(G)))
    return thr > np.asarray(qprec_diag_max)


def stats_guard_dims(T1_diag_max, cnt, MZcols, tol=I8_TOL):
    """Data dimensions of a statistics pass the check kernel flags: cnt_d * max_c zscale_c * 2^-55 > tol * max_i T1[d][ii].
    MZcols: [N][P] the <zz^T> columns that were digitised (their column maxima give the scales)."""
    sz = float(np.max(column_scales(MZcols))) * 2.0 ** -55
    return np.asarray(cnt) * sz > tol * np.asarray(T1_diag_max)
