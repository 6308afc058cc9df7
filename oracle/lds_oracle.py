"""numpy restatement (float64, batched over independent sequences) of the reference's VB smoother for a linear
dynamic system -- examples/Linear_Dynamic_System.py:47-76 of /root/reference.

TEST INFRASTRUCTURE ONLY: the checker of pyvb_b200's LDS kernels.  Pinned against fixtures produced by the literal
reference (tests/golden/lds_*.npz, oracle/gen_golden_lds.py) in tests/test_oracle_golden.py.

Model per sequence: columns A_i, C_i ~ N(0, (alpha0 I)^-1) under hstack; Q, R DiagonalGamma(a0, b0);
X_0 ~ N(0, I), X_t ~ N(A X_{t-1}, Q), Y_t ~ N(C X_t, R) observed.  One iteration (lines 69-76): X_t forwards, X_t
backwards, A columns, C columns, Q, R.  The arithmetic each update performs:
  X_t   Gaussian.update  src/pyvb/nodes/gaussian.py:102-123 with the messages of Multiplication.pass_up_m1_m2
        src/pyvb/nodes/node.py:203-227:  prec_t = Qbar (I at t = 0) + <A^T Qbar A> [t < T-1] + <C^T Rbar C>,
        mean = prec^-1 (Qbar A x_{t-1} + A^T Qbar x_{t+1} + C^T Rbar y_t)   
  A_i   hstack.pass_up_m1_m2  src/pyvb/nodes/nodes_todo.py:43-62  (Gauss-Seidel over the columns)
  Q, R  DiagonalGamma.update  src/pyvb/nodes/nodes_todo.py:187-190
State (B = number of sequences): A, Avar (B,q,q) [row k, column i]; C, Cvar (B,d,q); Qa, Qb (B,q); Ra, Rb (B,d);
X (B,T,q); Xcov (B,T,q,q).
"""
import numpy as np


class LDSOracle(object):
    KEYS = ("A", "Avar", "C", "Cvar", "Qa", "Qb", "Ra", "Rb", "X", "Xcov")

    def __init__(self, Y, q, alpha0=1e-3, a0=1e-3, b0=1e-3):
        Y = np.asarray(Y, dtype=np.float64)
        if Y.ndim == 2:
            Y = Y[None]
        self.Y = Y
        self.B, self.T, self.d = Y.shape
        self.q = int(q)
        self.alpha0, self.a0, self.b0 = float(alpha0), float(a0), float(b0)
        B, T, d, q = self.B, self.T, self.d, self.q
        self.A = np.zeros((B, q, q)); self.Avar = np.ones((B, q, q))
        self.C = np.zeros((B, d, q)); self.Cvar = np.ones((B, d, q))
        self.Qa = np.full((B, q), a0 + 0.5 * (T - 1)); self.Qb = np.ones((B, q))
        self.Ra = np.full((B, d), a0 + 0.5 * T); self.Rb = np.ones((B, d))
        self.X = np.zeros((B, T, q)); self.Xcov = np.tile(np.eye(q), (B, T, 1, 1))

    def load_state(self, st):
        for k in self.KEYS:
            if k in st:
                v = np.array(st[k], dtype=np.float64, copy=True)
                cur = getattr(self, k)
                setattr(self, k, v.reshape(cur.shape))

    def state(self):
        return {k: np.array(getattr(self, k), copy=True) for k in self.KEYS}

    # ------------------------------------------------------------------ pieces
    def _quad(self, M, Mvar, lam):
        """<M^T diag(lam) M>[i,j] = sum_k lam_k (M[k,i] M[k,j] + delta_ij Mvar[k,i])   (node.py:213-227)"""
        out = np.einsum("bk,bki,bkj->bij", lam, M, M)
        idx = np.arange(M.shape[2])
        out[:, idx, idx] += np.einsum("bk,bki->bi", lam, Mvar)
        return out

    def update_X(self):
        B, T, q = self.B, self.T, self.q
        Qbar, Rbar = self.Qa / self.Qb, self.Ra / self.Rb
        AQA = self._quad(self.A, self.Avar, Qbar)
        CRC = self._quad(self.C, self.Cvar, Rbar)
        eye = np.eye(q)[None]
        Qd = np.einsum("bk,kl->bkl", Qbar, np.eye(q))
        S0 = np.linalg.inv(eye + (AQA if T > 1 else 0.0) + CRC)
        Si = np.linalg.inv(Qd + AQA + CRC)
        ST = np.linalg.inv(Qd + CRC)
        cry = np.einsum("bki,bk,btk->bti", self.C, Rbar, self.Y)          # C^T Rbar y_t
        M1 = Qbar[:, :, None] * self.A                                      # Qbar A
        M2 = np.einsum("bki,bk->bik", self.A, Qbar)                         # A^T Qbar

        def one(t):
            rhs = cry[:, t].copy()
            if t > 0:
                rhs += np.einsum("bki,bi->bk", M1, self.X[:, t - 1])
            if t < T - 1:
                rhs += np.einsum("bik,bk->bi", M2, self.X[:, t + 1])
            S = S0 if t == 0 else (ST if t == T - 1 else Si)
            self.X[:, t] = np.einsum("bij,bj->bi", S, rhs)
            self.Xcov[:, t] = S

        for t in range(T):
            one(t)
        for t in range(T - 1, -1, -1):
            one(t)

    def _sums(self):
        EXX = np.einsum("bti,btj->btij", self.X, self.X) + self.Xcov
        SC = EXX.sum(1)
        SA = SC - EXX[:, -1]
        XX1 = np.einsum("btk,bti->bki", self.X[:, 1:], self.X[:, :-1])      # sum_t x_t,k x_{t-1},i
        YX = np.einsum("btk,bti->bki", self.Y, self.X)
        return EXX, SA, SC, XX1, YX

    @staticmethod
    def _columns(M, Mvar, lam, S, cross, alpha0):
        """Gauss-Seidel over the columns of an hstack (nodes_todo.py:53-62 + gaussian.py:117-123)."""
        q = M.shape[2]
        for i in range(q):
            prec = alpha0 + lam * S[:, i, i][:, None]                      # (B, rows)
            m2 = lam * cross[:, :, i]
            for j in range(q):
                if j != i:
                    m2 = m2 - lam * S[:, i, j][:, None] * M[:, :, j]
            M[:, :, i] = m2 / prec
            Mvar[:, :, i] = 1.0 / prec

    def update_params(self):
        EXX, SA, SC, XX1, YX = self._sums()
        Qbar, Rbar = self.Qa / self.Qb, self.Ra / self.Rb
        self._columns(self.A, self.Avar, Qbar, SA, XX1, self.alpha0)
        self._columns(self.C, self.Cvar, Rbar, SC, YX, self.alpha0)
        # DiagonalGamma.update (nodes_todo.py:187-190); children of Q are X_1 .. X_{T-1}, of R all Y_t
        q, d = self.q, self.d
        idx = np.arange(q)
        quadA = np.einsum("bki,bkj,bij->bk", self.A, self.A, SA) + np.einsum("bki,bii->bk", self.Avar, SA)
        xx = (SC - EXX[:, 0])[:, idx, idx]
        self.Qb = self.b0 + 0.5 * xx + 0.5 * quadA - np.einsum("bki,bki->bk", self.A, XX1)
        quadC = np.einsum("bki,bkj,bij->bk", self.C, self.C, SC) + np.einsum("bki,bii->bk", self.Cvar, SC)
        yy = np.einsum("btk,btk->bk", self.Y, self.Y)
        self.Rb = self.b0 + 0.5 * yy + 0.5 * quadC - np.einsum("bki,bki->bk", self.C, YX)
        self.Qa = np.full((self.B, q), self.a0 + 0.5 * (self.T - 1))
        self.Ra = np.full((self.B, d), self.a0 + 0.5 * self.T)

    def iterate(self):
        self.update_X()
        self.update_params()


def synth_lds(B, T, q, d, seed=0):
    """Independent synthetic sequences in the style of examples/Linear_Dynamic_System.py:20-44."""
    rng = np.random.RandomState(seed)
    Y = np.zeros((B, T, d))
    for b in range(B):
        A = rng.randn(q, q)
        A *= 0.9 / max(1e-9, np.max(np.abs(np.linalg.eigvals(A))))
        C = rng.randn(d, q) * 3.0
        rs, qs = np.sqrt(rng.rand(d) * 0.1), np.sqrt(rng.rand(q) * 0.1)
        x = rng.randn(q)
        for t in range(T):
            if t:
                x = A @ x + qs * rng.randn(q)
            Y[b, t] = C @ x + rs * rng.randn(d)
    return Y
