"""Graph -> plate compiler for the VB-PCA (missing data) pattern, and the node <-> engine binding.

The reference builds one Python object per data row (examples/PCA_missing_data.py:35-36) and pulls
messages recursively (nodes/gaussian.py:112-115).  Here the same graph is recognised once and
lowered to a single device-resident plate (pyvb_b200.engine.PlateEngine); every node keeps a binding
(kind, index) so that ``node.update()`` / ``node.qmu`` / ``node.qcov`` address the device state.

Recognised pattern (anything else raises NotImplementedError -- there is no CPU message passing):

    X_n ~ Gaussian(d, hstack(W_0..W_{q-1}) * Z_n + Mu, Beta)         n = 0..N-1
    W_i ~ Gaussian(d, 0, alpha*I | Gamma)   Mu ~ Gaussian(d, 0, alpha_mu*I)
    Z_n ~ Gaussian(q, m0, P0)               Beta ~ Gamma(d, a0, b0)
"""
import numpy as np

DEFAULT_MODE = "A"     # "A": reference-exact imputation semantics; "B": masked / marginalised
DEFAULT_ALGO = "auto"


def set_default_mode(mode):
    global DEFAULT_MODE
    assert mode in ("A", "B")
    DEFAULT_MODE = mode


def _crawl(start):
    """All nodes connected to `start` (same neighbourhood relation as Network.fetch_network)."""
    from . import nodes as nd
    seen, order, stack = set(), [], [start]
    while stack:
        n = stack.pop()
        if id(n) in seen:
            continue
        seen.add(id(n))
        order.append(n)
        nbrs = list(getattr(n, "children", []))
        if isinstance(n, nd.Gaussian):
            nbrs += [n.mean_parent, n.precision_parent]
        elif isinstance(n, (nd.Addition, nd.Multiplication)):
            nbrs += [n.A, n.B]
        elif isinstance(n, nd.hstack):
            nbrs += list(n.parents)
        stack.extend(nbrs)
    return order


def _scalar_times_eye(M):
    M = np.asarray(M)
    a = M[0, 0]
    if not np.array_equal(M, np.eye(M.shape[0]) * a):
        raise NotImplementedError("prior precision of W columns / Mu must be alpha*I (got a general matrix)")
    return float(a)


class PCAPlate(object):
    """One compiled VB-PCA plate.  kinds: 'W' (column i), 'Z' (row n), 'X' (row n), 'M', 'B', 'L' (ARD alpha i)."""

    ENGINE = None      # engine class; None = pyvb_b200.engine.PlateEngine (tests inject a recorder)

    def __init__(self, any_node, mode=None, algo=None):
        from . import nodes as nd
        PlateEngine = self.ENGINE
        if PlateEngine is None:
            from .engine import PlateEngine
        mode = DEFAULT_MODE if mode is None else mode
        algo = DEFAULT_ALGO if algo is None else algo
        allnodes = _crawl(any_node)
        hs = [n for n in allnodes if isinstance(n, nd.hstack)]
        fixed_z = False
        if len(hs) == 1:
            W = hs[0]
            Ws = list(W.parents)
        elif len(hs) == 0:
            # q = 1 without an hstack (src/tests.py:176-202, simple_PCA): W is ONE Gaussian column, every product is
            # Multiplication(W, z_n) with a scalar z_n (node.py:195-197, 205-207) -- the same plate with q = 1
            cols = [n for n in allnodes if isinstance(n, nd.Gaussian) and n.children
                    and all(isinstance(c, nd.Multiplication) and c.A is n for c in n.children)]
            # scalar regression (src/tests.py:100-128, simple_regression): y_n ~ N(x_n * A + B, noise) with CONSTANT 1 x 1
            # regressors x_n on the left: the same plate with d = q = 1, W := A, z_n := x_n fixed (never updated)
            regs = [n for n in allnodes if isinstance(n, nd.Gaussian) and n.shape == (1, 1) and n.children
                    and all(isinstance(c, nd.Multiplication) and c.B is n and isinstance(c.A, nd.Constant)
                            and c.A.shape == (1, 1) for c in n.children)]
            if len(cols) == 1:
                W = cols[0]
            elif len(regs) == 1 and not cols:
                W = regs[0]
                fixed_z = True
            else:
                raise NotImplementedError("expected one hstack (W), one Gaussian column multiplied by scalar z_n, or one "
                                          "scalar Gaussian multiplied by constant regressors")
            Ws = [W]
        else:
            raise NotImplementedError("expected exactly one hstack (W) in the graph, found %d" % len(hs))
        q = len(Ws)
        mults = list(W.children)
        if not mults or not all(isinstance(m, nd.Multiplication) and (m.B if fixed_z else m.A) is W for m in mults):
            raise NotImplementedError("children of W must be Multiplication(W, z_n) nodes")
        Zs, Xs, Mu, Beta = [], [], None, None
        for m in mults:
            z = m.A if fixed_z else m.B
            if fixed_z:
                if len(m.children) != 1:
                    raise NotImplementedError("each product x_n * A must have a single child")
            elif not isinstance(z, nd.Gaussian) or len(z.children) != 1 or len(m.children) != 1:
                raise NotImplementedError("each z_n must be a Gaussian with the single child W*z_n")
            add = m.children[0]
            if not isinstance(add, nd.Addition) or len(add.children) != 1:
                raise NotImplementedError("expected X_n ~ Gaussian(d, W*z_n + Mu, Beta)")
            other = add.B if add.A is m else add.A
            if not isinstance(other, nd.Gaussian):
                raise NotImplementedError("the offset Mu must be a Gaussian node")
            x = add.children[0]
            if not isinstance(x, nd.Gaussian) or x.children or x.mean_parent is not add:
                raise NotImplementedError("X_n must be a leaf Gaussian with mean W*z_n + Mu")
            if Mu is None:
                Mu, Beta = other, x.precision_parent
            if other is not Mu or x.precision_parent is not Beta:
                raise NotImplementedError("all rows must share Mu and the noise precision")
            Zs.append(z)
            Xs.append(x)
        if not isinstance(Beta, nd.Gamma):
            raise NotImplementedError("noise precision must be a Gamma node")
        d, N = W.shape[0], len(Xs)
        # priors
        ard = isinstance(Ws[0].precision_parent, nd.Gamma)
        Alphas = []
        alpha0 = 1e-3
        for w in Ws:
            if not isinstance(w.mean_parent, nd.Constant) or np.any(w.mean_parent.value != 0):
                raise NotImplementedError("W columns need a constant zero prior mean")
            if ard:
                if not isinstance(w.precision_parent, nd.Gamma) or len(w.precision_parent.children) != 1:
                    raise NotImplementedError("ARD: one Gamma per W column")
                Alphas.append(w.precision_parent)
            else:
                alpha0 = _scalar_times_eye(w.precision_parent.value)
        if ard and len(set((a.a0, a.b0) for a in Alphas)) != 1:
            raise NotImplementedError("ARD Gammas must share (a0, b0)")
        if not isinstance(Mu.mean_parent, nd.Constant) or np.any(Mu.mean_parent.value != 0):
            raise NotImplementedError("Mu needs a constant zero prior mean")
        alpha_mu = _scalar_times_eye(Mu.precision_parent.value)
        z0 = Zs[0]
        if fixed_z:
            m0, P0 = np.zeros((q, 1)), np.eye(q)          # (unused: the regressors are never updated)
            if not all(x.observed for x in Xs):
                raise NotImplementedError("regression pattern: every y_n must be observed")
        else:
            if not isinstance(z0.mean_parent, nd.Constant) or not isinstance(z0.precision_parent, nd.Constant):
                raise NotImplementedError("z_n needs constant prior mean and precision")
            m0, P0 = z0.mean_parent.value, z0.precision_parent.value
            for z in Zs:
                if not (np.array_equal(z.mean_parent.value, m0) and np.array_equal(z.precision_parent.value, P0)):
                    raise NotImplementedError("all z_n must share one prior")

        # data + initial state from the nodes (the reference's random init, gaussian.py:70-72)
        X = np.full((N, d), np.nan)
        for n, x in enumerate(Xs):
            if x.observed:
                X[n] = x._qmu[:, 0]
            elif x.partially_observed:
                X[n] = x.obs_value[:, 0]
        st = {
            "Wbar": np.hstack([w._qmu for w in Ws]),
            "Wvar": np.stack([np.diag(w._qcov) for w in Ws], 1),
            "mu": Mu._qmu[:, 0], "muvar": np.diag(Mu._qcov),
            "Zbar": np.stack([(z.value if fixed_z else z._qmu)[:, 0] for z in Zs]),
            "Sig": np.stack([np.zeros((q, q)) if fixed_z else z._qcov for z in Zs]),
            "Xhat": np.stack([x._qmu[:, 0] for x in Xs]),
            "V": np.stack([np.diag(x._qcov) for x in Xs]),
            "qb": Beta._qb,
        }
        if ard:
            st["al_qb"] = np.array([a._qb for a in Alphas])
        self.mode = mode
        self.engine = PlateEngine(X, q, mode=mode, alpha0=alpha0, alpha_mu=alpha_mu, a0=Beta.a0, b0=Beta.b0,
                                  ard=ard, ard_a0=Alphas[0].a0 if ard else 1e-3, ard_b0=Alphas[0].b0 if ard else 1e-3,
                                  P0=P0, m0=m0[:, 0], algo=algo)
        self.engine.set_state(st)
        self.Xdata = X
        self.N, self.d, self.q = N, d, q
        self.W, self.Ws, self.Mu, self.Beta, self.Zs, self.Xs, self.Alphas = W, Ws, Mu, Beta, Zs, Xs, Alphas
        self.index = {}
        for i, w in enumerate(Ws):
            self.index[id(w)] = ("W", i)
        self.fixed_z = fixed_z
        for n, z in enumerate(Zs):
            if not fixed_z:
                self.index[id(z)] = ("Z", n)
        for n, x in enumerate(Xs):
            self.index[id(x)] = ("X", n)
        self.index[id(Mu)] = ("M", 0)
        self.index[id(Beta)] = ("B", 0)
        for i, a in enumerate(Alphas):
            self.index[id(a)] = ("L", i)
        for n in Ws + ([] if fixed_z else Zs) + Xs + [Mu, Beta] + Alphas:
            n._binding = self
        self._host = None
        self._elbo_terms = None

    # ------------------------------------------------------------------ updates
    def _touch(self):
        self._host = None
        self._small = None
        self._rows = {}
        self._elbo_terms = None

    def run(self, kind, lo, hi):
        e = self.engine
        if kind == "W":
            e.update_W(lo, hi)
        elif kind == "Z":
            e.update_Z(lo, hi)
        elif kind == "X":
            e.update_X(lo, hi)
        elif kind == "M":
            e.update_Mu()
        elif kind == "B":
            e.update_Beta()
        elif kind == "L":
            e.update_Alpha(lo, hi)
        self._touch()

    def update(self, node):
        kind, i = self.index[id(node)]
        self.run(kind, i, i + 1)

    def schedule(self, nodes):
        """Coalesce an update order into batched plate operations [(kind, lo, hi)]."""
        out = []
        for n in nodes:
            key = self.index.get(id(n))
            if key is None:
                raise NotImplementedError("node %r is not part of the compiled plate" % (n,))
            kind, i = key
            if out and out[-1][0] == kind and out[-1][2] == i and kind in "WZXL":
                out[-1][2] = i + 1
            else:
                out.append([kind, i, i + 1])
        return [tuple(s) for s in out]

    def sweep(self, sched):
        for kind, lo, hi in sched:
            self.run(kind, lo, hi)

    def elbo(self):
        if self.fixed_z:
            raise NotImplementedError("the bound of the regression pattern is not evaluated (its regressors are constants)")
        v = self.engine.elbo()
        self._elbo_terms = None
        return v

    def log_lower_bound(self, node):
        """Per-node share of the bound.  The kernels evaluate the bound per node *group*; a group's total
        is apportioned equally over its members, so the sum over nodes (network.py:49) is exact."""
        from ._layout import GL_ELBO_W
        if self._elbo_terms is None:
            self.engine.elbo()
            g = self.engine.gl.cpu().numpy()
            self._elbo_terms = dict(zip("WMZXBL", g[GL_ELBO_W:GL_ELBO_W + 6]))
        kind, _ = self.index[id(node)]
        cnt = {"W": self.q, "M": 1, "Z": self.N, "X": self.N, "B": 1, "L": max(len(self.Alphas), 1)}[kind]
        return float(self._elbo_terms[kind]) / cnt

    # ------------------------------------------------------------------ state views
    def _state(self):
        if self._host is None:
            self.engine.check()
            self._host = self.engine.get_state()
        return self._host

    def _view(self, kind, i):
        """What one node's attributes need: the replicated state (O(D q)) for W / Mu / Beta / Alpha, plus ONE row of the plate
        for Z_i / X_i -- the manual update order of src/tests.py:312-316 reads an attribute after every single update, and a
        copy of the whole state each time would cost O(N q^2) per read."""
        if self._host is not None:
            return self._host
        if getattr(self, "_small", None) is None:
            self.engine.check()
            self._small = self.engine.get_state_small()
        if kind not in "ZX":
            return self._small
        rows = getattr(self, "_rows", None)
        if rows is None:
            rows = self._rows = {}
        if i not in rows:
            rows[i] = self.engine.get_row(i)
        r = rows[i]
        st = dict(self._small)
        for k, v in r.items():                                 # (indexable by i like the full state)
            st[k] = {i: v}
        return st

    def get(self, node, attr):
        kind, i = self.index[id(node)]
        st = self._view(kind, i)
        if attr == "qb":
            return st["qb"] if kind == "B" else float(st["al_qb"][i])
        if kind == "W":
            return st["Wbar"][:, i:i + 1] if attr == "qmu" else np.diag(st["Wvar"][:, i])
        if kind == "M":
            return st["mu"][:, None] if attr == "qmu" else np.diag(st["muvar"])
        if kind == "Z":
            return st["Zbar"][i][:, None] if attr == "qmu" else st["Sig"][i]
        if kind == "X":
            if self.mode == "A":
                return st["Xhat"][i][:, None] if attr == "qmu" else np.diag(st["V"][i])
            x = self.Xdata[i]
            miss = np.isnan(x)
            if attr == "qmu":
                pred = st["Wbar"] @ st["Zbar"][i] + st["mu"]
                return np.where(miss, pred, x)[:, None]
            return np.diag(np.where(miss, 1.0 / st["tau"], 0.0))
        raise AttributeError(attr)

    def set(self, node, attr, value):
        kind, i = self.index[id(node)]
        st = self._state()
        value = np.asarray(value, dtype=np.float64)
        key = {("W", "qmu"): "Wbar", ("W", "qcov"): "Wvar", ("M", "qmu"): "mu", ("M", "qcov"): "muvar",
               ("Z", "qmu"): "Zbar", ("Z", "qcov"): "Sig", ("X", "qmu"): "Xhat", ("X", "qcov"): "V",
               ("B", "qb"): "qb", ("L", "qb"): "al_qb"}[(kind, attr)]
        if kind == "W":
            st[key][:, i] = value[:, 0] if attr == "qmu" else np.diag(value)
        elif kind == "M":
            st[key][:] = value[:, 0] if attr == "qmu" else np.diag(value)
        elif kind == "Z":
            st[key][i] = value[:, 0] if attr == "qmu" else value
        elif kind == "X":
            if self.mode != "A":
                raise NotImplementedError("X nodes hold no free state in mode B")
            st[key][i] = value[:, 0] if attr == "qmu" else np.diag(value)
        elif kind == "B":
            st[key] = float(value)
        else:
            st[key][i] = float(value)
        self.engine.set_state(st)
        self._touch()


def bind(node, mode=None, algo=None):
    """Return the plate `node` belongs to, compiling the graph on first use: the VB-PCA pattern (PCAPlate) or the
    linear-dynamic-system pattern of the reference's LDS scripts (lds_plate.LDSPlate)."""
    b = getattr(node, "_binding", None)
    if b is None:
        try:
            b = PCAPlate(node, mode=mode, algo=algo)
        except NotImplementedError as pca_err:
            from .lds_plate import LDSPlate
            try:
                b = LDSPlate(node)
            except NotImplementedError as lds_err:
                raise NotImplementedError("the graph matches no compiled pattern -- VB-PCA: %s; LDS: %s (there is no CPU "
                                          "message passing)" % (pca_err, lds_err))
    return b
