"""TEST INFRASTRUCTURE (never imported by pyvb_b200/): lane-level numpy restatement of the blocked symmetric sweep
(pyvb_b200/csrc/kernels_k2s.cu), the kernel that replaces cho_factor / cho_solve(., I) / dot(qcov, .) / q_ln_det of
Gaussian.update (/root/reference/src/pyvb/nodes/gaussian.py:117-123).

It follows the kernel step by step ON THE PACKED ROW as it lies in HBM / shared memory (lower triangle, element (i, j) at
i (i + 1) / 2 + j; eta behind it): the 32 lanes of a warp are numpy vectors, a DMMA.8x8x4 is emulated from its fragment layout
(lane l: a = A[l/4][l%4], b = B[l%4][l/4], c0/c1 = C[l/4][2 (l%4) + {0,1}]).  Per tile column K: the 8 x 8 pivot tiles of the
warp's matrices factored together, M_KK = L D L^T and X = L^-1 by elimination on [M_KK | I] (lane (m, r) = row r of matrix m's
tile, one published column entry per lane and pivot plus the pivot lane's row of X), kept in the tile's own slots (X below the
diagonal, 1 / d on it); the old panel tiles as A fragments, W_J = old_J X^T on the emulated tensor core, through the panel slots
into A fragments, the new panel (W_J D^-1) X in the symmetric sweep convention (above the pivot tile into the slots of the
transposed elements), the pivot tile -X^T D^-1 X, the trailing update M_IJ -= (W_I D^-1) W_J^T of the lower tiles; then
zbar = Sigma eta (lane = row) and <zz^T> = Sigma + zbar zbar^T tile by tile, from the state -Sigma.  A wrong fragment index, a
missing mirror read or a read after an in-place write shows up here, on the CPU.  (One rounding per fused multiply-add is not
modelled.)"""
import numpy as np

LANE = np.arange(32)
GID, QD = LANE >> 2, LANE & 3


def tri(i):
    return i * (i + 1) // 2


def mz_pitch(q):
    p = ((tri(q) + 7) & ~7) + q + 1
    while p % 8 != 4:
        p += 1
    return p


def dmma(c0, c1, a, b):
    """D(8x8) += A(8x4) B(4x8) from the per-lane fragments"""
    A = a.reshape(8, 4)                       # lane = 4 row + k
    B = b.reshape(8, 4).T                     # lane = 4 n + k  ->  B[k][n]
    C = A @ B
    return c0 + C[GID, 2 * QD], c1 + C[GID, 2 * QD + 1]


def pivot_tile(stg, q, K, mpw, pr, pos):
    """8 x 8 pivot tiles K of the mpw matrices: M_KK = L D L^T by elimination without pivoting, X = L^-1 by the same eliminations
    on the identity; lane (m, r) = row r of matrix m's tile (with mpw < 4 the other lanes shadow a lane of the same row).
    Out, in the slots of the tile's lower triangle: X below the diagonal (its diagonal is 1), 1 / d on the diagonal."""
    pitch = mz_pitch(q)
    c0 = 8 * K
    m, r = (LANE >> 3) % mpw, LANE & 7
    base = m * pitch
    i = c0 + r
    a = np.empty((32, 8))
    for j in range(8):
        a[:, j] = np.where(j <= r, stg[base + tri(i) + c0 + np.minimum(j, r)], stg[base + tri(c0 + j) + i])
    x = np.zeros((32, 8))
    dinv = np.ones(32)
    for k in range(8):
        bc = np.full((mpw, 8), np.nan)
        bc[m, r] = a[:, k]                                        # column k of the reduced tile (rows >= k are current) = its row k
        xk = np.zeros((mpw, 8))
        w = r == k
        xk[m[w]] = x[w]                                           # the pivot lane publishes its row of X (entries < k)
        B, XK = bc[m], xk[m]
        rc = 1.0 / B[:, k]
        dinv = np.where(r == k, rc, dinv)
        l = np.where(r > k, a[:, k] * rc, 0.0)
        for j in range(k + 1, 8):
            a[:, j] = a[:, j] - l * B[:, j]
        for j in range(k):
            x[:, j] = x[:, j] - l * XK[:, j]
        x[:, k] = -l
    act = (LANE >> 3) < mpw
    pr *= np.where(act, dinv, 1.0)
    pos &= np.where(act, dinv > 0.0, True)
    for j in range(8):
        w = j < r
        stg[(base + tri(i) + c0 + j)[w]] = x[:, j][w]
    stg[base + tri(i) + c0 + r] = dinv


def xfrag(st, c0, rr, cc):
    """element (rr, cc) of the unit lower triangular X kept below the diagonal of pivot tile c0 (rr, cc local lane vectors)"""
    v = st[tri(c0 + np.maximum(rr, cc)) + c0 + np.minimum(rr, cc)]
    return np.where(rr > cc, v, np.where(rr == cc, 1.0, 0.0))


def sweep_tile(stg, q, K, mpw, pr, pos):
    """tile column K of the mpw matrices of a group; stg = their packed rows [mpw * pitch] (modified in place)"""
    pitch, nbt = mz_pitch(q), q // 8
    c0 = 8 * K
    pivot_tile(stg, q, K, mpw, pr, pos)
    for m in range(mpw):
        st = stg[m * pitch:(m + 1) * pitch]
        of = {}
        for J in range(nbt):
            for h in range(2):
                if J != K:
                    of[J, h] = (st[tri(8 * J + GID) + c0 + 4 * h + QD] if J > K else st[tri(c0 + 4 * h + QD) + 8 * J + GID]).copy()
        xa = [xfrag(st, c0, GID, 4 * h + QD) for h in range(2)]           # X[gid][4h + qd]: A fragment of X = B fragment of X^T
        xb = [xfrag(st, c0, 4 * h + QD, GID) for h in range(2)]           # X[4h + qd][gid]: B fragment of X = A fragment of X^T
        dk = [st[tri(c0 + 4 * h + QD) + c0 + 4 * h + QD].copy() for h in range(2)]   # 1 / d_k, k = 4h + qd
        # W_J = old_J X^T
        wt = {}
        for J in range(nbt):
            if J != K:
                c = dmma(np.zeros(32), np.zeros(32), of[J, 0], xa[0])
                wt[J] = dmma(c[0], c[1], of[J, 1], xa[1])
        # the pivot tile: -inv(M_KK) = -X^T D^-1 X
        c = dmma(np.zeros(32), np.zeros(32), xb[0], xb[0] * dk[0])
        pv = dmma(c[0], c[1], xb[1], xb[1] * dk[1])
        for e in range(2):
            w = 2 * QD + e <= GID
            st[(tri(c0 + GID) + c0 + 2 * QD + e)[w]] = -pv[e][w]
        # W into the panel slots (for the change of layout), back as A fragments
        for J in range(nbt):
            if J == K:
                continue
            if J > K:
                o = tri(8 * J + GID) + c0 + 2 * QD
                st[o] = wt[J][0]
                st[o + 1] = wt[J][1]
            else:
                st[tri(c0 + 2 * QD) + 8 * J + GID] = wt[J][0]
                st[tri(c0 + 2 * QD + 1) + 8 * J + GID] = wt[J][1]
        wa, wd = {}, {}
        for I in range(nbt):
            for h in range(2):
                if I != K:
                    wa[I, h] = (st[tri(8 * I + GID) + c0 + 4 * h + QD] if I > K else st[tri(c0 + 4 * h + QD) + 8 * I + GID]).copy()
                    wd[I, h] = wa[I, h] * dk[h]
        # T_J = (W_J D^-1) X into the panel slots (sweep convention: +M_JK inv(M_KK))
        for J in range(nbt):
            if J == K:
                continue
            c = dmma(np.zeros(32), np.zeros(32), wd[J, 0], xb[0])
            tt = dmma(c[0], c[1], wd[J, 1], xb[1])
            if J > K:
                o = tri(8 * J + GID) + c0 + 2 * QD
                st[o] = tt[0]
                st[o + 1] = tt[1]
            else:
                st[tri(c0 + 2 * QD) + 8 * J + GID] = tt[0]
                st[tri(c0 + 2 * QD + 1) + 8 * J + GID] = tt[1]
        # trailing update: M_IJ -= (W_I D^-1) W_J^T, lower tiles outside row / column K
        for I in range(nbt):
            for J in range(I + 1):
                if I == K or J == K:
                    continue
                o = tri(8 * I + GID) + 8 * J + 2 * QD
                c0v, c1v = st[o].copy(), st[o + 1].copy()         # (upper half of a diagonal tile: junk, never stored)
                c0v, c1v = dmma(c0v, c1v, -wd[I, 0], wa[J, 0])
                c0v, c1v = dmma(c0v, c1v, -wd[I, 1], wa[J, 1])
                w0 = np.ones(32, bool) if I > J else (2 * QD <= GID)
                w1 = np.ones(32, bool) if I > J else (2 * QD + 1 <= GID)
                st[o[w0]] = c0v[w0]
                st[(o + 1)[w1]] = c1v[w1]


def sweep_group(stg, q, mpw):
    """One group: stg = mpw packed rows (modified in place: <zz^T> packed | zbar).  Returns Sigma packed [mpw, P], logdet [mpw]."""
    pitch, P, nbt = mz_pitch(q), tri(q), q // 8
    PP = (P + 7) & ~7
    pr, pos = np.ones(32), np.ones(32, bool)
    for K in range(nbt):
        sweep_tile(stg, q, K, mpw, pr, pos)
    with np.errstate(invalid="ignore", divide="ignore"):
        lg = np.where(pos, np.log(pr), np.nan)
    logdet = (-0.5 * lg.reshape(4, 8).sum(axis=1))[:mpw]
    Sigma = np.stack([-stg[m * pitch:m * pitch + P] for m in range(mpw)])
    npass = (mpw * q + 31) // 32
    z = []
    for p in range(npass):
        idx = p * 32 + LANE
        m, i = np.minimum(idx // q, mpw - 1), idx % q
        acc = np.zeros(32)
        for j in range(q):
            v = stg[m * pitch + tri(np.maximum(i, j)) + np.minimum(i, j)]
            acc = acc + v * stg[m * pitch + PP + j]
        z.append(-acc)
    for p in range(npass):
        idx = p * 32 + LANE
        w = idx // q < mpw
        m, i = (idx // q)[w], (idx % q)[w]
        stg[m * pitch + PP + i] = z[p][w]
    for m in range(mpw):
        st = stg[m * pitch:(m + 1) * pitch]
        for I in range(nbt):
            zi = st[PP + 8 * I + GID].copy()
            for J in range(I + 1):
                for e in range(2):
                    zj = st[PP + 8 * J + 2 * QD + e]
                    o = tri(8 * I + GID) + 8 * J + 2 * QD + e
                    w = np.ones(32, bool) if I > J else (2 * QD + e <= GID)
                    st[o[w]] = (zi * zj - st[o])[w]
    return Sigma, logdet


def sweep_solve(A, eta, mpw=4):
    """A [N, q, q] SPD, eta [N, q] -> Sigma = A^-1 [N, q, q], zbar [N, q], ln prod diag chol(A) [N], <zz^T> [N, q, q]
    through the packed-row, lane-level restatement above (N is padded to whole groups with identity matrices)."""
    A = np.asarray(A, dtype=np.float64)
    N, q, _ = A.shape
    pitch, P = mz_pitch(q), tri(q)
    PP = (P + 7) & ~7
    ii, jj = np.tril_indices(q)
    Sg, Zb, Ld, M2 = np.empty((N, q, q)), np.empty((N, q)), np.empty(N), np.empty((N, q, q))
    for n0 in range(0, N, mpw):
        stg = np.full(mpw * pitch, 7.0)                           # (pads hold junk the kernel must not depend on)
        for m in range(mpw):
            n = n0 + m
            stg[m * pitch:m * pitch + P] = A[n][ii, jj] if n < N else np.eye(q)[ii, jj]
            stg[m * pitch + PP:m * pitch + PP + q] = eta[n] if n < N else 0.0
        Sp, ld = sweep_group(stg, q, mpw)
        for m in range(min(mpw, N - n0)):
            n = n0 + m
            S = np.zeros((q, q))
            S[ii, jj] = Sp[m]
            Sg[n] = S + np.tril(S, -1).T
            Zb[n] = stg[m * pitch + PP:m * pitch + PP + q]
            S[ii, jj] = stg[m * pitch:m * pitch + P]
            M2[n] = S + np.tril(S, -1).T
            Ld[n] = ld[m]
    return Sg, Zb, Ld, M2
