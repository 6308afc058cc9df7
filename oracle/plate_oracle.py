"""CPU restatement ("plate oracle") of pyvb's VB-PCA-with-missing-data path.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline legs may import this file; nothing under
``pyvb_b200/`` does.  It is the checker, never the product.

What it restates (reference file:line, all under /root/reference/src/pyvb):
  * ``Gaussian.update``               nodes/gaussian.py:102-134
  * ``Gaussian.log_lower_bound``      nodes/gaussian.py:136-151
  * ``Gaussian.pass_up_m1_m2``        nodes/gaussian.py:179-183
  * ``Addition.pass_up_m1_m2``        nodes/node.py:95-110
  * ``Multiplication.pass_up_m1_m2``  nodes/node.py:182-232  (the tr(<w_i w_j^T> Lambda) contraction, 213-227)
  * ``Multiplication.pass_down_ExxT`` nodes/node.py:244-276
  * ``hstack.pass_up_m1_m2``          nodes/nodes_todo.py:43-62
  * ``Gamma.update/log_lower_bound``  nodes/nodes_todo.py:113-157
  * ``Network.learn`` sweep order     network.py:40-96
for the model of examples/PCA_missing_data.py:31-37, one row of the data
matrix per "plate" element, vectorised over rows with numpy.

Two modes (findings 3-5):
  mode "A"  reference-exact *imputation*: a partially observed X_n is a latent
            Gaussian conditioned on its observed entries; every message is
            unmasked.  Pinned against the literal reference (tests/golden).
  mode "B"  masked / marginalised: Lambda_n = tau*diag(mask_n).  The Z, W-column
            and Mu updates are pinned against the reference's generic operators
            driven with a per-row Constant precision (tests/golden/modeB_*.npz).
            The Gamma update and the ELBO under masking are NOT expressible in
            the reference => "parity unpinned" for those two pieces; they are
            the reference formulas restricted to observed entries and coincide
            with mode A (hence with the reference) when nothing is missing.

Logs: the reference computes ``log(det(.))`` / ``log(prod(diag(chol)))``
literally (node.py:302, gaussian.py:120), which under/overflows for D >= 108.
This restatement uses sums of logs; it equals the reference wherever the
reference is finite.  All the reference's ELBO quirks are kept on purpose
(``q_ln_det = .5/log(prod diag chol)`` is a *division*, gaussian.py:120; the
sign of the partial-row entropy, gaussian.py:150; ``ln<tau>`` instead of
``<ln tau>``, nodes_todo.py:144-147).
"""
import numpy as np
from scipy import special

LN2PI = np.log(2.0 * np.pi)


def tril_index(q):
    """Packed lower-triangular index p(i,j) = i(i+1)/2 + j, i >= j."""
    ii, jj = np.tril_indices(q)
    return ii, jj


def pack_sym(M):
    """(..., q, q) symmetric -> (..., P) packed lower triangle."""
    q = M.shape[-1]
    ii, jj = tril_index(q)
    return M[..., ii, jj]


def unpack_sym(Pk, q):
    ii, jj = tril_index(q)
    out = np.zeros(Pk.shape[:-1] + (q, q), dtype=Pk.dtype)
    out[..., ii, jj] = Pk
    out[..., jj, ii] = Pk
    return out


class PlateOracle(object):
    """State + updates of the VB-PCA plate model.

    State arrays (float64):
      Xhat (N,D)  <X_n>; NaN = missing (mode B only; mode A never holds NaN)
      V    (N,D)  diag of Cov(X_n) (mode A imputation variances; 0 in mode B)
      Wbar (D,q), Wvar (D,q)   column means and diag of the column covariances
      mu (D,), muvar (D,)
      Zbar (N,q), Sig (N,q,q)
      qa, qb                   noise Gamma
      al_qa (q,), al_qb (q,)   ARD Gammas (ard=True) -- else alpha0 constant
      qldW (q,), qldMu, qldZ (N,), qldX (N,)   the reference's ``q_ln_det``
    """

    def __init__(self, X, q, mode="B", alpha0=1e-3, alpha_mu=1e-3, a0=1e-3, b0=1e-3,
                 ard=False, ard_a0=1e-3, ard_b0=1e-3, P0=None, m0=None):
        X = np.asarray(X, dtype=np.float64)
        self.mode = mode
        self.N, self.D = X.shape
        self.q = q
        self.X = X.copy()
        self.O = ~np.isnan(X)
        nobs = self.O.sum(1)
        self.full = nobs == self.D
        self.latent = nobs == 0
        self.partial = ~(self.full | self.latent)
        self.alpha0 = float(alpha0)
        self.alpha_mu = float(alpha_mu)
        self.a0, self.b0 = float(a0), float(b0)
        self.ard = ard
        self.ard_a0, self.ard_b0 = float(ard_a0), float(ard_b0)
        self.P0 = np.eye(q) if P0 is None else np.asarray(P0, dtype=np.float64)
        self.m0 = np.zeros(q) if m0 is None else np.asarray(m0, dtype=np.float64).reshape(q)
        N, D = self.N, self.D
        # deterministic stand-in init; parity runs overwrite these
        self.Wbar = np.zeros((D, q))
        self.Wvar = np.ones((D, q))
        self.mu = np.zeros(D)
        self.muvar = np.ones(D)
        self.Zbar = np.zeros((N, q))
        self.Sig = np.tile(np.eye(q), (N, 1, 1))
        if mode == "A":
            self.Xhat = np.where(self.O, X, 0.0)
            self.V = np.where(self.O, 0.0, 1.0)
            self.qa = self.a0 + 0.5 * N * D          # nodes_todo.py:125-128 counts every child
        else:
            self.Xhat = X.copy()
            self.V = np.zeros((N, D))
            self.qa = self.a0 + 0.5 * self.O.sum()
        self.qb = 0.5
        self.al_qa = np.full(q, self.ard_a0 + 0.5 * D)
        self.al_qb = np.ones(q)
        self.qldW = np.zeros(q)
        self.qldMu = 0.0
        self.qldZ = np.zeros(N)
        self.qldX = np.zeros(N)

    # ------------------------------------------------------------------ helpers
    @property
    def tau(self):
        return self.qa / self.qb

    def alpha(self):
        return self.al_qa / self.al_qb if self.ard else np.full(self.q, self.alpha0)

    def E(self):
        """Effective observation mask of the messages (all ones in mode A)."""
        if self.mode == "A":
            return np.ones((self.N, self.D))
        return self.O.astype(np.float64)

    def Xe(self):
        """<X> with zeros where the effective mask is 0."""
        return np.where(np.isnan(self.Xhat), 0.0, self.Xhat)

    def M2(self):
        """<z z^T>_n  (gaussian.py:162-168)."""
        return self.Zbar[:, :, None] * self.Zbar[:, None, :] + self.Sig

    def G(self):
        """G_d = wbar_d wbar_d^T + diag_i Cov_i[d,d]   (node.py:219-224)."""
        G = self.Wbar[:, :, None] * self.Wbar[:, None, :]
        idx = np.arange(self.q)
        G[:, idx, idx] += self.Wvar
        return G

    # ------------------------------------------------------------------ updates
    def update_W_col(self, i):
        """hstack.pass_up_m1_m2 (nodes_todo.py:43-62) + Gaussian.update (gaussian.py:117-123)."""
        tau, E, q = self.tau, self.E(), self.q
        M2 = self.M2()
        T1i = E.T @ M2[:, i, :]                                  # (D,q): sum_n E_nd <zz^T>_n[i,j]
        prec = self.alpha()[i] + tau * T1i[:, i]
        R = E * (self.Xe() - self.mu[None, :])
        m2 = tau * (R.T @ self.Zbar[:, i])
        for j in range(q):
            if j != i:
                m2 -= tau * T1i[:, j] * self.Wbar[:, j]          # uses CURRENT wbar_j (Gauss-Seidel)
        self.Wbar[:, i] = m2 / prec
        self.Wvar[:, i] = 1.0 / prec
        self.qldW[i] = 0.5 / (0.5 * np.sum(np.log(prec)))

    def update_W(self):
        """All columns in sequence; the row sums are shared (one GEMM) but every column still sees the
        already-updated columns j < i, exactly as calling update_W_col(i) for i = 0..q-1."""
        tau, E, q = self.tau, self.E(), self.q
        T1 = (E.T @ self.M2().reshape(self.N, q * q)).reshape(self.D, q, q)
        R = E * (self.Xe() - self.mu[None, :])
        T2 = R.T @ self.Zbar
        al = self.alpha()
        for i in range(q):
            prec = al[i] + tau * T1[:, i, i]
            m2 = tau * T2[:, i]
            for j in range(q):
                if j != i:
                    m2 -= tau * T1[:, i, j] * self.Wbar[:, j]
            self.Wbar[:, i] = m2 / prec
            self.Wvar[:, i] = 1.0 / prec
            self.qldW[i] = 0.5 / (0.5 * np.sum(np.log(prec)))

    def update_Z(self, lo=0, hi=None):
        """Multiplication.pass_up_m1_m2 requester=z (node.py:203-227) + Gaussian.update."""
        hi = self.N if hi is None else hi
        tau, q = self.tau, self.q
        E = self.E()[lo:hi]
        Gf = self.G().reshape(self.D, q * q)
        L = self.P0[None] + tau * (E @ Gf).reshape(-1, q, q)     # K1
        R = E * (self.Xe()[lo:hi] - self.mu[None, :])
        eta = (self.P0 @ self.m0)[None, :] + tau * (R @ self.Wbar)
        U = np.linalg.cholesky(L)                                # K2
        Sig = np.linalg.inv(L)
        Sig = 0.5 * (Sig + np.transpose(Sig, (0, 2, 1)))
        self.Sig[lo:hi] = Sig
        self.Zbar[lo:hi] = np.einsum("nij,nj->ni", Sig, eta)
        logprod = np.sum(np.log(np.diagonal(U, axis1=1, axis2=2)), axis=1)
        self.qldZ[lo:hi] = 0.5 / logprod

    def update_X(self, lo=0, hi=None):
        """Mode A imputation (gaussian.py:102-134) for rows that are not fully observed."""
        if self.mode != "A":
            return
        hi = self.N if hi is None else hi
        tau = self.tau
        sl = slice(lo, hi)
        upd = ~self.full[sl]
        m = self.Zbar[sl] @ self.Wbar.T + self.mu[None, :]
        O = self.O[sl]
        xh = np.where(O, np.where(O, self.X[sl], 0.0), m)
        v = np.where(O, 0.0, 1.0 / tau)
        self.Xhat[sl] = np.where(upd[:, None], xh, self.Xhat[sl])
        self.V[sl] = np.where(upd[:, None], v, self.V[sl])
        self.qldX[sl] = np.where(upd, 0.5 / (0.5 * self.D * np.log(tau)), self.qldX[sl])

    def update_Mu(self):
        """Addition.pass_up_m1_m2 requester=Mu (node.py:105-109) + Gaussian.update."""
        tau, E = self.tau, self.E()
        prec = self.alpha_mu + tau * E.sum(0)
        resid = E * (self.Xe() - self.Zbar @ self.Wbar.T)
        self.mu = tau * resid.sum(0) / prec
        self.muvar = 1.0 / prec
        self.qldMu = 0.5 / (0.5 * np.sum(np.log(prec)))

    def _resid2(self):
        """sum over effective entries of <(x_nd - w_d.z_n - mu_d)^2>."""
        E, q = self.E(), self.q
        m = self.Zbar @ self.Wbar.T + self.mu[None, :]
        Gf = self.G().reshape(self.D, q * q)
        M2f = self.M2().reshape(self.N, q * q)
        trGM = M2f @ Gf.T                                         # tr(G_d <zz^T>_n)
        wz = self.Zbar @ self.Wbar.T
        per = (self.Xe() - m) ** 2 + self.V + trGM - wz ** 2 + self.muvar[None, :]
        return float(np.sum(E * per))

    def update_Beta(self):
        """Gamma.update (nodes_todo.py:130-138)."""
        self.qb = self.b0 + 0.5 * self._resid2()

    def update_Alpha(self):
        """Gamma.update for one ARD precision per W column."""
        if not self.ard:
            return
        self.al_qb = self.ard_b0 + 0.5 * (np.sum(self.Wbar ** 2, 0) + np.sum(self.Wvar, 0))

    # ------------------------------------------------------------------ ELBO
    @staticmethod
    def _gamma_llb(a0, b0, qa, qb):
        """Gamma.log_lower_bound (nodes_todo.py:149-157)."""
        Elnx = special.digamma(qa) - np.log(qb)
        ret = (a0 - 1) * Elnx - special.gammaln(a0) + a0 * np.log(b0) - b0 * (qa / qb)
        ret -= (qa - 1) * Elnx - special.gammaln(qa) + qa * np.log(qb) - qb * (qa / qb)
        return ret

    def elbo_terms(self):
        D, q, N = self.D, self.q, self.N
        t = {}
        al = self.alpha()
        lndet_al = D * (np.log(self.al_qa) - np.log(self.al_qb)) if self.ard else np.full(q, D * np.log(self.alpha0))
        ww = np.sum(self.Wbar ** 2, 0) + np.sum(self.Wvar, 0)
        t["W"] = float(np.sum(-0.5 * D * LN2PI + 0.5 * lndet_al - 0.5 * al * ww
                              - (-0.5 * D * LN2PI - 0.5 * self.qldW - 0.5 * D)))
        mm = np.sum(self.mu ** 2) + np.sum(self.muvar)
        t["Mu"] = float(-0.5 * D * LN2PI + 0.5 * D * np.log(self.alpha_mu) - 0.5 * self.alpha_mu * mm
                        - (-0.5 * D * LN2PI - 0.5 * self.qldMu - 0.5 * D))
        # Z_n: general constant prior N(m0, P0^-1)
        M2 = self.M2()
        lndetP0 = np.linalg.slogdet(self.P0)[1]
        trPM = np.einsum("ij,nji->n", self.P0, M2)
        quad = trPM + self.m0 @ self.P0 @ self.m0 - 2.0 * (self.Zbar @ (self.P0 @ self.m0))
        t["Z"] = float(np.sum(-0.5 * q * LN2PI + 0.5 * lndetP0 - 0.5 * quad
                              - (-0.5 * q * LN2PI - 0.5 * self.qldZ - 0.5 * q)))
        # X_n
        tau = self.tau
        nE = float(self.E().sum())
        x = -0.5 * nE * LN2PI + 0.5 * nE * (np.log(self.qa) - np.log(self.qb)) - 0.5 * tau * self._resid2()
        if self.mode == "A":
            lat = self.latent
            x -= np.sum(-0.5 * D * LN2PI - 0.5 * self.qldX[lat] - 0.5 * D)
            par = self.partial
            nmiss = (~self.O[par]).sum(1)
            lnv = np.where(self.O[par], 0.0, np.log(np.where(self.O[par], 1.0, self.V[par]))).sum(1)
            x -= np.sum(0.5 * nmiss * LN2PI - 0.5 * lnv - 0.5 * nmiss)
        t["X"] = float(x)
        t["Beta"] = float(self._gamma_llb(self.a0, self.b0, self.qa, self.qb))
        t["Alpha"] = float(np.sum(self._gamma_llb(self.ard_a0, self.ard_b0, self.al_qa, self.al_qb))) if self.ard else 0.0
        return t

    def elbo(self):
        return float(sum(self.elbo_terms().values()))

    # ------------------------------------------------------------------ sweep
    def iterate(self):
        """One sweep in the order Network.fetch_network induces for the shipped
        script (network.py:58-96): W_0..W_{q-1}, Z_0..Z_{N-1},
        [Alpha_*], X_0, Mu, X_1..X_{N-1}, Beta.  Returns the ELBO."""
        self.update_W()
        self.update_Z()
        self.update_Alpha()
        self.update_X(0, 1)
        self.update_Mu()
        self.update_X(1, self.N)
        self.update_Beta()
        return self.elbo()

    def learn(self, niters):
        return [self.iterate() for _ in range(niters)]

    # ------------------------------------------------------------------ state io
    STATE_KEYS = ("Wbar", "Wvar", "mu", "muvar", "Zbar", "Sig", "Xhat", "V", "qb", "al_qb")

    def state(self):
        return {k: np.array(getattr(self, k), dtype=np.float64, copy=True) for k in self.STATE_KEYS}

    def load_state(self, st):
        for k in self.STATE_KEYS:
            if k in st:
                v = np.array(st[k], dtype=np.float64, copy=True)
                setattr(self, k, float(v) if k == "qb" else v)


def synth_pca(N, D, q, missing, seed=0, noise_prec=20.0):
    """Synthetic workload generalising examples/PCA_missing_data.py:16-27 with
    i.i.d. Bernoulli(missing) erasures."""
    rng = np.random.RandomState(seed)
    W = rng.randn(D, q)
    Z = rng.randn(N, q)
    mu = rng.randn(D)
    X = Z @ W.T + mu[None, :] + rng.randn(N, D) * np.sqrt(1.0 / noise_prec)
    if missing > 0:
        X[rng.rand(N, D) < missing] = np.nan
    return X
