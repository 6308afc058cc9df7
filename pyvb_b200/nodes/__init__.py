"""Node classes mirroring pyvb.nodes (reference: src/pyvb/nodes/{node,gaussian,nodes_todo}.py).

Same names, constructor signatures, attributes and error conventions as the reference, so that
examples/PCA_missing_data.py:31-45 runs unmodified.  The objects only *describe* the graph and hold
the random initial state (drawn from the global numpy stream in the reference's order, so a seeded
script gets the reference's initialisation).  All updates run on the GPU: the first
``update()`` / ``log_lower_bound()`` / ``Network.learn()`` compiles the graph into a plate
(``pyvb_b200.plate``) backed by the CUDA engine, after which ``qmu``/``qcov``/``qa``/``qb`` are
views of device state.  Graphs outside the compiled patterns raise NotImplementedError -- there is
no CPU message-passing fallback for ``update()``.

The message getters (``pass_up_m1_m2``, ``pass_down_Ex/ExxT``) are an *inspection* API: they evaluate
ONE node's message on the host from the device-backed state, with the reference's semantics
(node.py:95-129, 182-276; nodes_todo.py:43-62; gaussian.py:179-183).  The update kernels never call
them -- inside the plate the same quantities are fused sums over all rows.
"""
import numpy as np

from .. import plate as _plate


class ConjugacyError(ValueError):
    """nodes_todo.py:8-10"""

    def __init__(self, message):
        ValueError.__init__(self, message)


class Node(object):
    """node.py:6-50"""

    def __init__(self, shape):
        self.children = []
        self.shape = shape

    def addChild(self, child):
        self.children.append(child)

    def update(self):
        pass

    def log_lower_bound(self):
        return 0.

    def __add__(self, other):
        return Addition(self, other)

    def __mul__(self, other):
        return Multiplication(self, other)

    def __rmul__(self, other):
        return Multiplication(other, self)


class Constant(Node):
    """node.py:279-311"""

    def __init__(self, value):
        Node.__init__(self, value.shape)
        self.shape = value.shape
        self.value = value

    def pass_down_Ex(self):
        return self.value

    def pass_down_ExxT(self):
        return np.dot(self.value, self.value.T)

    def pass_down_ExTx(self):
        return np.dot(self.value.T, self.value)

    def pass_down_lndet(self):
        # stable restatement of log(det(value)) (node.py:302 underflows for D >= 108)
        return np.linalg.slogdet(self.value)[1]


class Addition(Node):
    """node.py:52-129 (graph description only; messages are fused into the CUDA kernels)"""

    def __init__(self, A, B):
        assert A.shape == B.shape, "Bad shapes for addition"
        Node.__init__(self, A.shape)
        self.A = Constant(A) if type(A) == np.ndarray else A
        self.B = Constant(B) if type(B) == np.ndarray else B
        self.A.addChild(self)
        self.B.addChild(self)

    def pass_down_Ex(self):
        return self.A.pass_down_Ex() + self.B.pass_down_Ex()

    def pass_down_ExxT(self):
        """<(A+B)(A+B)^T> for independent A, B (node.py:120-129)"""
        cross = np.dot(self.A.pass_down_Ex(), self.B.pass_down_Ex().T)
        return self.A.pass_down_ExxT() + self.B.pass_down_ExxT() + cross + cross.T

    def pass_up_m1_m2(self, requester):
        """Children's (m1, m2) with the co-parent's mean taken out of m2 (node.py:95-110)."""
        m1, m2 = _sum_child_messages(self)
        other = self.B if requester is self.A else self.A
        return m1, m2 - np.dot(m1, other.pass_down_Ex())


def _sum_child_messages(node):
    msgs = [c.pass_up_m1_m2(node) for c in node.children]
    return sum(m[0] for m in msgs), sum(m[1] for m in msgs)


class Multiplication(Node):
    """node.py:131-276 (graph description only)"""

    def __init__(self, A, B):
        m1, n1 = A.shape
        m2, n2 = B.shape
        assert n1 == m2, "incompatible multiplication dimensions"
        assert n2 == 1, "right hand object must be a vector"
        Node.__init__(self, (m1, n2))
        self.A = Constant(A) if type(A) == np.ndarray else A
        self.B = Constant(B) if type(B) == np.ndarray else B
        self.A.addChild(self)
        self.B.addChild(self)

    def pass_down_Ex(self):
        return np.dot(self.A.pass_down_Ex(), self.B.pass_down_Ex())

    def _second_moment_of_A(self):
        """<a_i a_j^T> for the columns of an hstack A as a (q, q, d, d) array: outer products of the column means plus the
        column covariances on the i == j blocks (node.py:213-224)."""
        Abar = self.A.pass_down_Ex()
        G = np.einsum("ai,bj->ijab", Abar, Abar)
        for i, col in enumerate(self.A.parents):
            G[i, i] += col.qcov
        return G

    def pass_down_ExxT(self):
        """<(AB)(AB)^T> (node.py:244-276): column A times scalar B, Constant matrix A, or hstack A."""
        BBt = self.B.pass_down_ExxT()
        if self.A.shape[1] == 1:
            return self.A.pass_down_ExxT() * float(BBt[0, 0])
        if isinstance(self.A, Constant):
            return np.dot(self.A.value, np.dot(BBt, self.A.value.T))
        if hasattr(self.A, "parents"):
            return np.einsum("ijab,ij->ab", self._second_moment_of_A(), BBt)
        raise NotImplementedError("pass_down_ExxT for this left operand")

    def pass_up_m1_m2(self, requester):
        """node.py:182-232.  To an hstack A the raw 4-tuple (sum m1, sum m2, <B>, <BB^T>) is forwarded; to B the contraction
        m1 = tr(<a_i a_j^T> sum m1), m2 = <A>^T sum m2 -- the quantity K1 evaluates for all rows at once."""
        m1s, m2s = _sum_child_messages(self)
        if requester is self.A:
            if self.A.shape[1] == 1:
                return m1s * float(self.B.pass_down_ExxT()[0, 0]), float(self.B.pass_down_Ex()[0, 0]) * m2s
            if hasattr(self.A, "parents"):
                return m1s, m2s, self.B.pass_down_Ex(), self.B.pass_down_ExxT()
            raise NotImplementedError("messages to this left operand")
        m2 = np.dot(self.A.pass_down_Ex().T, m2s)
        if self.A.shape[1] == 1:
            return np.trace(np.dot(self.A.pass_down_ExxT(), m1s)), m2
        if isinstance(self.A, Constant):
            return np.dot(self.A.value.T, np.dot(m1s, self.A.value)), m2
        if hasattr(self.A, "parents"):
            return np.einsum("ijab,ba->ij", self._second_moment_of_A(), m1s), m2
        raise NotImplementedError("messages through this left operand")


class hstack(Node):
    """nodes_todo.py:12-62: a matrix whose columns are Gaussian nodes."""

    def __init__(self, parents):
        dims = [e.shape[0] for e in parents]
        shape = (dims[0], len(parents))
        Node.__init__(self, shape)
        assert type(parents) == list
        assert np.all(dims[0] == np.array(dims)), "dimensions incompatible"
        self.parents = parents
        self.shape = shape
        [e.addChild(self) for e in self.parents]

    def pass_down_Ex(self):
        return np.hstack([e.pass_down_Ex() for e in self.parents])

    def pass_down_ExxT(self):
        return np.sum([p.pass_down_ExxT() for p in self.parents], 0)

    def pass_down_ExTx(self):
        raise NotImplementedError

    def pass_up_m1_m2(self, requester):
        """Message to column i (nodes_todo.py:43-62): m1 = sum_n m1_n <b b^T>_n[i, i],
        m2 = sum_n (m2_n <b_i>_n - sum_{j != i} m1_n <b b^T>_n[i, j] <a_j>) -- what K3 + the W update evaluate."""
        msgs = [c.pass_up_m1_m2(self) for c in self.children]
        if self.shape[1] == 1:
            return sum(m[0] for m in msgs), sum(m[1] for m in msgs)
        i = self.parents.index(requester)
        m1 = sum(m[0] * float(m[3][i, i]) for m in msgs)
        m2 = sum(m[1] * float(m[2][i, 0]) for m in msgs)
        for j, col in enumerate(self.parents):
            if j != i:
                m2 = m2 - np.dot(sum(m[0] * float(m[3][i, j]) for m in msgs), col.pass_down_Ex())
        return m1, m2


class Gamma(object):
    """nodes_todo.py:88-157: isotropic precision (noise, or ARD precision of a W column)."""

    def __init__(self, dim, a0, b0):
        self.shape = (dim, dim)
        self.a0 = a0
        self.b0 = b0
        self.children = []
        self._binding = None
        self.update_a()
        self._qb = np.random.rand()          # same draw as nodes_todo.py:119

    def addChild(self, child):
        self.children.append(child)
        self.update_a()

    def update_a(self):
        self.qa = self.a0
        for child in self.children:
            self.qa += 0.5 * child.shape[0]

    @property
    def qb(self):
        if self._binding is not None:
            return self._binding.get(self, "qb")
        return self._qb

    @qb.setter
    def qb(self, v):
        if self._binding is not None:
            self._binding.set(self, "qb", v)
        else:
            self._qb = v

    def update(self):
        _plate.bind(self).update(self)

    def log_lower_bound(self):
        return _plate.bind(self).log_lower_bound(self)

    def pass_down_Ex(self):
        return np.eye(self.shape[0]) * self.qa / self.qb

    def pass_down_lndet(self):
        return self.shape[0] * (np.log(self.qa) - np.log(self.qb))


class Gaussian(Node):
    """gaussian.py:9-183"""

    def __init__(self, dim, pmu, pprec):
        Node.__init__(self, (dim, 1))
        assert pmu.shape == self.shape, "Parent node (or array) has incorrect dimension"
        if type(pmu) == np.ndarray:
            self.mean_parent = Constant(pmu)
        elif isinstance(pmu, (Gaussian, Addition, Multiplication, Constant)):
            self.mean_parent = pmu
        else:
            raise ConjugacyError("mean parent for a Gaussian node should be one of:\nGaussian\nConstant\nAddition"
                                 "\nMultiplication\nnumpy array. \n\n" + str(type(pmu)) + " is invalid")
        assert pprec.shape == (self.shape[0], self.shape[0]), "Parent precision array has incorrect dimension"
        if type(pprec) == np.ndarray:
            self.precision_parent = Constant(pprec)
        elif isinstance(pprec, (Gamma, DiagonalGamma, Wishart, Constant)):
            self.precision_parent = pprec
        else:
            raise ConjugacyError("Precision parent for a Gaussian node should be one of:\nGamma\nDiagonalGamma"
                                 "\nWishart\nConstant\nnumpy array. \n\n" + str(type(pprec)) + " is invalid")
        self.mean_parent.addChild(self)
        self.precision_parent.addChild(self)
        self.observed = False
        self.partially_observed = False
        self._binding = None
        # random initial solution: the same three draws, in the same order, as gaussian.py:70-72
        self._qmu = np.random.randn(self.shape[0], 1)
        self.qprec = np.eye(self.shape[0]) * np.random.rand()
        self._qcov = np.linalg.inv(self.qprec)

    # ---- state: local until the graph is compiled, device views afterwards
    @property
    def qmu(self):
        if self._binding is not None:
            return self._binding.get(self, "qmu")
        return self._qmu

    @qmu.setter
    def qmu(self, v):
        if self._binding is not None:
            self._binding.set(self, "qmu", v)
        else:
            self._qmu = v

    @property
    def qcov(self):
        if self._binding is not None:
            return self._binding.get(self, "qcov")
        return self._qcov

    @qcov.setter
    def qcov(self, v):
        if self._binding is not None:
            self._binding.set(self, "qcov", v)
        else:
            self._qcov = v

    def observe(self, val):
        """gaussian.py:74-100; NaN entries of val are missing data."""
        assert val.shape == self.shape, "Bad shape for observation data"
        if self._binding is not None:
            raise NotImplementedError("observe() after the graph has been compiled to a plate")
        if np.isnan(val).all():
            return
        elif np.isnan(val).any():
            self.partially_observed = True
            self.obs_value = val
            self.obs_index = np.nonzero(1 - np.isnan(val))[0]
            self.missing_index = np.nonzero(np.isnan(val))[0]
        else:
            self.observed = True
            self._qmu = val
            self._qcov = np.zeros(self._qcov.shape)

    def update(self):
        if self.observed:
            return
        _plate.bind(self).update(self)

    def log_lower_bound(self):
        return _plate.bind(self).log_lower_bound(self)

    def pass_down_Ex(self):
        return self.qmu

    def pass_down_ExxT(self):
        return np.dot(self.qmu, self.qmu.T) + self.qcov

    def pass_down_ExTx(self):
        return np.trace(self.pass_down_ExxT())

    def pass_up_m1_m2(self, requester):
        """(<Lambda>, <Lambda> qmu), unmasked even for a partially observed node (gaussian.py:179-183)."""
        pp = self.precision_parent.pass_down_Ex()
        return pp, np.dot(pp, self.qmu)


def _out_of_scope(name, ref):
    def __init__(self, *a, **k):
        raise NotImplementedError("%s (%s) is outside the VB-PCA hot path built here; see DESIGN.md" % (name, ref))
    return type(name, (object,), {"__init__": __init__, "__doc__": "out of scope: " + ref})


class DiagonalGamma(object):
    """nodes_todo.py:159-204: independent Gamma precisions for the entries of a Gaussian (Q and R of the LDS scripts)."""

    def __init__(self, dim, a0s, b0s):
        self.shape = (dim, dim)
        assert a0s.size == self.shape[0]
        assert b0s.size == self.shape[0]
        self.a0s = a0s.flatten()
        self.b0s = b0s.flatten()
        self.children = []
        self._binding = None
        self.update_a()
        self._qb = np.random.rand()          # ONE scalar, the same draw as nodes_todo.py:176

    def addChild(self, child):
        assert child.shape == (self.shape[0], 1)
        self.children.append(child)
        self.update_a()

    def update_a(self):
        self._qa = self.a0s + 0.5 * len(self.children)

    @property
    def qa(self):
        return self._qa

    @property
    def qb(self):
        if self._binding is not None:
            return self._binding.get(self, "qb")
        return self._qb

    def update(self):
        _plate.bind(self).update(self)

    def pass_down_Ex(self):
        return np.diag(self.qa / self.qb)

    def pass_down_lndet(self):
        return np.log(np.prod(self.qa / self.qb))

    def log_lower_bound(self):
        """Sum over the entries of the Gamma terms (nodes_todo.py:198-203); host arithmetic on 2 dim device-backed scalars."""
        from scipy import special
        qa, qb = self.qa, self.qb * np.ones(self.shape[0])
        elnx = special.digamma(qa) - np.log(qb)
        prior = (self.a0s - 1) * elnx - special.gammaln(self.a0s) + self.a0s * np.log(self.b0s) - self.b0s * qa / qb
        entropy = (qa - 1) * elnx - special.gammaln(qa) + qa * np.log(qb) - qa
        return float(np.sum(prior - entropy))


DiagonalGaussian = _out_of_scope("DiagonalGaussian", "gaussian.py:185-203")
Wishart = _out_of_scope("Wishart", "nodes_todo.py:205-234")
Transpose = _out_of_scope("Transpose", "nodes_todo.py:65-82")
