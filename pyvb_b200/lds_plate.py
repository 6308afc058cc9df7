"""Graph -> batched-smoother compiler for the linear-dynamic-system pattern of the reference's LDS scripts
(/root/reference/examples/Linear_Dynamic_System.py:47-76, examples/LDS_knowns_in_A.py:50-86):

    A = hstack(A_0..A_{q-1}),  C = hstack(C_0..C_{q-1})            columns ~ Gaussian(0, alpha*I)
    Q = DiagonalGamma(q, a0s, b0s),  R = DiagonalGamma(d, a0s, b0s)
    X_0 ~ Gaussian(q, 0, I),  X_t ~ Gaussian(q, A * X_{t-1}, Q),  Y_t ~ Gaussian(d, C * X_t, R)  observed
    [A_i.observe(column with NaN = unknown)]                        known entries of A

The reference updates one node at a time.  Here the node-level ``update()`` calls are RECORDED and executed lazily: when a
value is read (``qmu``, ``qcov``, ``qb``, ``pass_down_Ex`` ...) the recorded calls must spell whole iterations of the
scripts' sweep -- all X_t forwards, all X_t backwards, the A columns, the C columns, Q, R -- and k such iterations become
ONE launch of the batched smoother kernel (pyvb_lds_iterate_known_f64, niters = k).  Any other update order raises
NotImplementedError: there is no CPU message passing.
"""
import numpy as np


class LDSPlate(object):
    ENGINE = None      # engine class; None = pyvb_b200.lds.LDSEngine (tests inject a recorder)

    def __init__(self, any_node):
        from . import nodes as nd
        from .plate import _crawl, _scalar_times_eye
        allnodes = _crawl(any_node)
        hs = [n for n in allnodes if isinstance(n, nd.hstack)]
        if len(hs) != 2:
            raise NotImplementedError("LDS pattern: expected two hstacks (A and C), found %d" % len(hs))

        def is_transition(h):
            return all(isinstance(m, nd.Multiplication) and m.A is h and len(m.children) == 1
                       and isinstance(m.children[0], nd.Gaussian) and not m.children[0].observed for m in h.children)
        trans = [h for h in hs if h.children and is_transition(h)]
        if len(trans) != 1:
            raise NotImplementedError("LDS pattern: one hstack must be the state transition A (A * X_t is the mean of X_t+1)")
        A = trans[0]
        C = hs[0] if hs[1] is A else hs[1]
        q, d = A.shape[1], C.shape[0]
        if A.shape != (q, q) or C.shape[1] != q:
            raise NotImplementedError("LDS pattern: A must be q x q and C d x q")
        # the chain: X_0 is the state that is nobody's transition child
        nxt = {}
        for m in A.children:
            nxt[id(m.B)] = m.children[0]
        firsts = [m.B for m in A.children if not isinstance(m.B.mean_parent, nd.Multiplication)]
        if len(firsts) != 1:
            raise NotImplementedError("LDS pattern: a single chain X_0 -> X_1 -> ... is required")
        Xs = [firsts[0]]
        while id(Xs[-1]) in nxt:
            Xs.append(nxt[id(Xs[-1])])
        T = len(Xs)
        if T != len(A.children) + 1:
            raise NotImplementedError("LDS pattern: the transitions do not form one chain")
        x0 = Xs[0]
        if not (isinstance(x0.mean_parent, nd.Constant) and not np.any(x0.mean_parent.value)
                and isinstance(x0.precision_parent, nd.Constant) and np.array_equal(x0.precision_parent.value, np.eye(q))):
            raise NotImplementedError("LDS pattern: X_0 ~ N(0, I)")
        Q = Xs[1].precision_parent
        # observations: every X_t has exactly one C * X_t child with an observed Y_t
        emis = {id(m.B): m for m in C.children}
        Y = np.zeros((T, d))
        R = None
        for t, x in enumerate(Xs):
            m = emis.get(id(x))
            if m is None or len(m.children) != 1 or not m.children[0].observed:
                raise NotImplementedError("LDS pattern: every state needs one fully observed Y_t ~ N(C X_t, R)")
            y = m.children[0]
            Y[t] = y._qmu[:, 0]
            R = y.precision_parent if R is None else R
            if y.precision_parent is not R or (t > 0 and x.precision_parent is not Q):
                raise NotImplementedError("LDS pattern: shared Q and R")
        if not isinstance(Q, nd.DiagonalGamma) or not isinstance(R, nd.DiagonalGamma):
            raise NotImplementedError("LDS pattern: Q and R must be DiagonalGamma nodes")
        As, Cs = list(A.parents), list(C.parents)
        alpha = set()
        for col in As + Cs:
            if not isinstance(col.mean_parent, nd.Constant) or np.any(col.mean_parent.value):
                raise NotImplementedError("LDS pattern: the columns of A and C need a constant zero prior mean")
            alpha.add(_scalar_times_eye(col.precision_parent.value))
        if len(alpha) != 1:
            raise NotImplementedError("LDS pattern: one prior precision alpha*I for all columns of A and C")
        for g in (Q, R):
            if len(set(g.a0s.tolist())) != 1 or len(set(g.b0s.tolist())) != 1 or g.a0s[0] != Q.a0s[0] or g.b0s[0] != Q.b0s[0]:
                raise NotImplementedError("LDS pattern: one (a0, b0) for all entries of Q and R")
        if any(c.partially_observed or c.observed for c in Cs) or any(a.observed for a in As):
            raise NotImplementedError("LDS pattern: only entries of A may be known (LDS_knowns_in_A.py)")
        known = None
        if any(a.partially_observed for a in As):
            known = np.full((q, q), np.nan)
            for i, a in enumerate(As):
                if a.partially_observed:
                    known[:, i] = a.obs_value[:, 0]
        Engine = self.ENGINE
        if Engine is None:
            from .lds import LDSEngine as Engine
        self.engine = Engine(Y[None], q, alpha0=alpha.pop(), a0=float(Q.a0s[0]), b0=float(Q.b0s[0]), A_known=known)
        # the nodes' random initial state (gaussian.py:70-72, nodes_todo.py:176: qb starts as ONE random scalar)
        self.engine.set_state({
            "A": np.hstack([a._qmu for a in As])[None], "Avar": np.stack([np.diag(a._qcov) for a in As], 1)[None],
            "C": np.hstack([c._qmu for c in Cs])[None], "Cvar": np.stack([np.diag(c._qcov) for c in Cs], 1)[None],
            "Qb": np.full((1, q), float(np.asarray(Q._qb).reshape(-1)[0])) if np.ndim(Q._qb) == 0 else np.asarray(Q._qb)[None],
            "Rb": np.full((1, d), float(np.asarray(R._qb).reshape(-1)[0])) if np.ndim(R._qb) == 0 else np.asarray(R._qb)[None],
            "X": np.stack([x._qmu[:, 0] for x in Xs])[None]})
        self.T, self.q, self.d = T, q, d
        self.index = {}
        for t, x in enumerate(Xs):
            self.index[id(x)] = ("X", t)
        for i, a in enumerate(As):
            self.index[id(a)] = ("A", i)
        for i, c in enumerate(Cs):
            self.index[id(c)] = ("C", i)
        self.index[id(Q)] = ("Q", 0)
        self.index[id(R)] = ("R", 0)
        for n in Xs + As + Cs + [Q, R]:
            n._binding = self
        # one iteration of the scripts' sweep (Linear_Dynamic_System.py:69-76)
        self.sweep = ([("X", t) for t in range(T)] + [("X", t) for t in range(T - 1, -1, -1)] + [("A", i) for i in range(q)]
                      + [("C", i) for i in range(q)] + [("Q", 0), ("R", 0)])
        self.pending = []
        self.iterations = 0
        self.launches = 0
        self._host = None

    # ------------------------------------------------------------------ updates (recorded, executed lazily)
    def update(self, node):
        self.pending.append(self.index[id(node)])
        self._host = None

    def flush(self):
        if not self.pending:
            return
        n = len(self.sweep)
        k, rem = divmod(len(self.pending), n)
        if rem or self.pending != self.sweep * k:
            self.pending = []
            raise NotImplementedError(
                "only whole iterations of the LDS sweep (all X_t forwards, all X_t backwards, the A columns, the C columns, "
                "Q, R: examples/Linear_Dynamic_System.py:69-76) are compiled; there is no CPU message passing")
        self.pending = []
        self.engine.iterate(k)
        self.engine.check()
        self.iterations += k
        self.launches += 1

    def schedule(self, nodes):
        raise NotImplementedError("Network.learn on the LDS pattern: the reference's LDS scripts drive the sweep by hand "
                                  "(Linear_Dynamic_System.py:69-76); call update() on the nodes in that order")

    def log_lower_bound(self, node):
        raise NotImplementedError("the bound of the LDS model is not evaluated (the reference's LDS scripts never ask for it)")

    # ------------------------------------------------------------------ state views
    def _state(self):
        self.flush()
        if self._host is None:
            self._host = self.engine.get_state()
        return self._host

    def get(self, node, attr):
        kind, i = self.index[id(node)]
        st = self._state()
        if kind == "X":
            return st["X"][0, i][:, None] if attr == "qmu" else st["Xcov"][0, i]
        if kind in "AC":
            m, v = st[kind][0], st[kind + "var"][0]
            return m[:, i:i + 1] if attr == "qmu" else np.diag(v[:, i])
        if attr == "qb":
            return st[kind + "b"][0]
        if attr == "qa":
            return st[kind + "a"][0]
        raise AttributeError(attr)

    def set(self, node, attr, value):
        raise NotImplementedError("the state of a compiled LDS plate is read-only from the node API")
