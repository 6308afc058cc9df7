"""Node classes mirroring pyvb.nodes (reference: src/pyvb/nodes/{node,gaussian,nodes_todo}.py).

Same names, constructor signatures, attributes and error conventions as the reference, so that
examples/PCA_missing_data.py:31-45 runs unmodified.  The objects only *describe* the graph and hold
the random initial state (drawn from the global numpy stream in the reference's order, so a seeded
script gets the reference's initialisation).  All arithmetic runs on the GPU: the first
``update()`` / ``log_lower_bound()`` / ``Network.learn()`` compiles the graph into a plate
(``pyvb_b200.plate``) backed by the CUDA engine, after which ``qmu``/``qcov``/``qa``/``qb`` are
views of device state.  Graphs outside the VB-PCA pattern raise NotImplementedError -- there is
no CPU message-passing fallback.
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
        raise NotImplementedError("messages are fused into the CUDA kernels (see DESIGN.md)")


def _out_of_scope(name, ref):
    def __init__(self, *a, **k):
        raise NotImplementedError("%s (%s) is outside the VB-PCA hot path built here; see DESIGN.md" % (name, ref))
    return type(name, (object,), {"__init__": __init__, "__doc__": "out of scope: " + ref})


DiagonalGaussian = _out_of_scope("DiagonalGaussian", "gaussian.py:185-203")
DiagonalGamma = _out_of_scope("DiagonalGamma", "nodes_todo.py:159-204")
Wishart = _out_of_scope("Wishart", "nodes_todo.py:205-234")
Transpose = _out_of_scope("Transpose", "nodes_todo.py:65-82")
