"""TEST INFRASTRUCTURE (never imported by pyvb_b200/): lane-level numpy restatement of the blocked symmetric sweep on the
TILE-SWIZZLED working layout (pyvb_b200/csrc/kernels_k2s.cu: zsolve_tsweep_kernel), the kernel that replaces cho_factor /
cho_solve(., I) / dot(qcov, .) / q_ln_det of Gaussian.update (/root/reference/src/pyvb/nodes/gaussian.py:117-123).

Same algorithm as oracle/sweep_oracle.py (factored pivot tiles, W = old X^T, M_IJ -= (W D^-1) W^T, ...), other data movement: the
packed row is re-laid-out in place into swizzled 8 x 8 tiles (everything read before anything is written), the diagonal tiles are
kept whole, zbar = Sigma eta runs on the emulated tensor core (eta in column 0 of the B fragment), and the <zz^T> pass puts the row
back into the packed order.  The 32 lanes of a warp are numpy vectors; a wrong swizzle, fragment index or a read after an in-place
write shows up here, on the CPU."""
import numpy as np
LANE = np.arange(32)
GID, QD = LANE >> 2, LANE & 3
def tri(i): return i * (i + 1) // 2
def mz_pitch(q):
    p = ((tri(q) + 7) & ~7) + q + 1
    while p % 8 != 4: p += 1
    return p
def sw(r, c):
    return r * 8 + ((((c >> 2) ^ ((r >> 1) & 1))) << 2) + (c & 3)
def toff(I, J): return (tri(I) + J) * 64
def dmma(c0, c1, a, b):
    A = a.reshape(8, 4); B = b.reshape(8, 4).T; C = A @ B
    return c0 + C[GID, 2 * QD], c1 + C[GID, 2 * QD + 1]
def ts_of(q):
    nbt = q // 8
    return tri(nbt) * 64 + q

def to_tiles(st, q):
    """packed row [P | pad | eta] in st[0:pitch] -> tile layout in st[0:ts] (all reads first, then all writes)"""
    P, nbt = tri(q), q // 8
    PP = (P + 7) & ~7
    regs = {}
    for I in range(nbt):
        for J in range(I + 1):
            for e in range(2):
                row, col = 8 * I + GID, 8 * J + 2 * QD + e
                hi, lo = np.maximum(row, col), np.minimum(row, col)
                regs[I, J, e] = st[tri(hi) + lo].copy()
    eta = [st[np.minimum(PP + 32 * k + LANE, PP + q - 1)].copy() for k in range((q + 31) // 32)]
    for I in range(nbt):
        for J in range(I + 1):
            o = toff(I, J) + sw(GID, 2 * QD)
            st[o] = regs[I, J, 0]; st[o + 1] = regs[I, J, 1]
    E = tri(nbt) * 64
    for k in range((q + 31) // 32):
        w = 32 * k + LANE < q
        st[(E + 32 * k + LANE)[w]] = eta[k][w]

def pivot_tile(stg, q, K, mpw, pr, pos):
    ts = ts_of(q)
    m, r = (LANE >> 3) % mpw, LANE & 7
    base = m * ts + toff(K, K)
    a = np.empty((32, 8)); x = np.zeros((32, 8)); dinv = np.ones(32)
    for j in range(8): a[:, j] = stg[base + sw(r, j)]
    for k in range(8):
        bc = np.full((mpw, 8), np.nan); bc[m, r] = a[:, k]
        xk = np.zeros((mpw, 8)); w = r == k; xk[m[w]] = x[w]
        B, XK = bc[m], xk[m]
        rc = 1.0 / B[:, k]
        dinv = np.where(r == k, rc, dinv)
        l = np.where(r > k, a[:, k] * rc, 0.0)
        for j in range(k + 1, 8): a[:, j] = a[:, j] - l * B[:, j]
        for j in range(k): x[:, j] = x[:, j] - l * XK[:, j]
        x[:, k] = -l
    act = (LANE >> 3) < mpw
    pr *= np.where(act, dinv, 1.0); pos &= np.where(act, dinv > 0.0, True)
    for j in range(8):
        w = j < r
        stg[(base + sw(r, j))[w]] = x[:, j][w]
    stg[base + sw(r, r)] = dinv

def afrag(st, I, K, h):
    """A fragment (gid, 4h + qd) of block (I, K) of the symmetric state, I != K"""
    return (st[toff(I, K) + sw(GID, 4 * h + QD)] if I > K else st[toff(K, I) + sw(4 * h + QD, GID)]).copy()
def cstore(st, J, K, v0, v1):
    """accumulator-layout values of block (J, K), J != K"""
    if J > K:
        o = toff(J, K) + sw(GID, 2 * QD); st[o] = v0; st[o + 1] = v1
    else:
        st[toff(K, J) + sw(2 * QD, GID)] = v0; st[toff(K, J) + sw(2 * QD + 1, GID)] = v1

def sweep_tile(stg, q, K, mpw, pr, pos):
    ts, nbt = ts_of(q), q // 8
    pivot_tile(stg, q, K, mpw, pr, pos)
    for m in range(mpw):
        st = stg[m * ts:(m + 1) * ts]
        pt = toff(K, K)
        of = {(J, h): afrag(st, J, K, h) for J in range(nbt) if J != K for h in range(2)}
        xa, xb, dk = [], [], []
        for h in range(2):
            rr, cc = GID, 4 * h + QD
            v = st[pt + sw(rr, cc)]; xa.append(np.where(rr > cc, v, np.where(rr == cc, 1.0, 0.0)))
            rr, cc = 4 * h + QD, GID
            v = st[pt + sw(rr, cc)]; xb.append(np.where(rr > cc, v, np.where(rr == cc, 1.0, 0.0)))
            dk.append(st[pt + sw(4 * h + QD, 4 * h + QD)].copy())
        wt = {}
        for J in range(nbt):
            if J != K:
                c = dmma(np.zeros(32), np.zeros(32), of[J, 0], xa[0]); wt[J] = dmma(c[0], c[1], of[J, 1], xa[1])
        c = dmma(np.zeros(32), np.zeros(32), xb[0], xb[0] * dk[0]); pv = dmma(c[0], c[1], xb[1], xb[1] * dk[1])
        o = pt + sw(GID, 2 * QD); st[o] = -pv[0]; st[o + 1] = -pv[1]
        for J in range(nbt):
            if J != K: cstore(st, J, K, wt[J][0], wt[J][1])
        wa = {(I, h): afrag(st, I, K, h) for I in range(nbt) if I != K for h in range(2)}
        wd = {k: v * dk[k[1]] for k, v in wa.items()}
        for J in range(nbt):
            if J == K: continue
            c = dmma(np.zeros(32), np.zeros(32), wd[J, 0], xb[0]); tt = dmma(c[0], c[1], wd[J, 1], xb[1])
            cstore(st, J, K, tt[0], tt[1])
        for I in range(nbt):
            for J in range(I + 1):
                if I == K or J == K: continue
                o = toff(I, J) + sw(GID, 2 * QD)
                c0v, c1v = st[o].copy(), st[o + 1].copy()
                c0v, c1v = dmma(c0v, c1v, -wd[I, 0], wa[J, 0]); c0v, c1v = dmma(c0v, c1v, -wd[I, 1], wa[J, 1])
                st[o] = c0v; st[o + 1] = c1v

def group(stg, q, mpw):
    """stg = mpw stage slots of ts doubles, each holding a packed row in its first pitch doubles"""
    ts, P, nbt = ts_of(q), tri(q), q // 8
    PP = (P + 7) & ~7; E = tri(nbt) * 64
    for m in range(mpw): to_tiles(stg[m * ts:(m + 1) * ts], q)
    pr, pos = np.ones(32), np.ones(32, bool)
    for K in range(nbt): sweep_tile(stg, q, K, mpw, pr, pos)
    with np.errstate(invalid="ignore", divide="ignore"):
        lg = np.where(pos, np.log(pr), np.nan)
    logdet = (-0.5 * lg.reshape(4, 8).sum(axis=1))[:mpw]
    Sig = []
    for m in range(mpw):
        st = stg[m * ts:(m + 1) * ts]
        # z = Sigma eta on the tensor core: eta in column 0 of the B fragment
        z = {}
        for I in range(nbt):
            z0, z1 = np.zeros(32), np.zeros(32)
            for J in range(nbt):
                for h in range(2):
                    if I == J: a = st[toff(I, I) + sw(GID, 4 * h + QD)].copy()
                    else: a = afrag(st, I, J, h)
                    b = np.where(GID == 0, st[E + 8 * J + 4 * h + QD], 0.0)
                    z0, z1 = dmma(z0, z1, a, b)
            z[I] = -z0                                   # valid on the lanes with qd == 0: z[8 I + gid]
        for I in range(nbt):
            w = QD == 0
            st[(E + 8 * I + GID)[w]] = z[I][w]           # (after every lane has read eta)
        out, sg = {}, {}
        for I in range(nbt):
            zi = st[E + 8 * I + GID].copy()
            for J in range(I + 1):
                o = toff(I, J) + sw(GID, 2 * QD)
                for e in range(2):
                    c = st[o + e].copy(); zj = st[E + 8 * J + 2 * QD + e]
                    out[I, J, e] = zi * zj - c; sg[I, J, e] = -c
        zb = [st[np.minimum(E + 32 * k + LANE, E + q - 1)].copy() for k in range((q + 31) // 32)]
        S = np.zeros(P)
        for I in range(nbt):
            for J in range(I + 1):
                for e in range(2):
                    row, col = 8 * I + GID, 8 * J + 2 * QD + e
                    w = col <= row
                    st[(tri(row) + col)[w]] = out[I, J, e][w]; S[(tri(row) + col)[w]] = sg[I, J, e][w]
        for k in range((q + 31) // 32):
            w = 32 * k + LANE < q
            st[(PP + 32 * k + LANE)[w]] = zb[k][w]
        Sig.append(S)
    return np.stack(Sig), logdet

def tsweep_solve(A, eta, mpw=4):
    A = np.asarray(A, dtype=np.float64); N, q, _ = A.shape
    pitch, P, ts = mz_pitch(q), tri(q), ts_of(q)
    PP = (P + 7) & ~7; ii, jj = np.tril_indices(q)
    Sg, Zb, Ld, M2 = np.empty((N, q, q)), np.empty((N, q)), np.empty(N), np.empty((N, q, q))
    for n0 in range(0, N, mpw):
        stg = np.full(mpw * ts, 7.0)
        for m in range(mpw):
            n = n0 + m
            stg[m * ts:m * ts + P] = A[n][ii, jj] if n < N else np.eye(q)[ii, jj]
            stg[m * ts + PP:m * ts + PP + q] = eta[n] if n < N else 0.0
        Sp, ld = group(stg, q, mpw)
        for m in range(min(mpw, N - n0)):
            n = n0 + m; S = np.zeros((q, q)); S[ii, jj] = Sp[m]; Sg[n] = S + np.tril(S, -1).T
            Zb[n] = stg[m * ts + PP:m * ts + PP + q]
            S[ii, jj] = stg[m * ts:m * ts + P]; M2[n] = S + np.tril(S, -1).T; Ld[n] = ld[m]
    return Sg, Zb, Ld, M2
