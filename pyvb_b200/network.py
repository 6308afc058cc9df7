"""Network: node registry, graph crawl, learn loop (reference: src/pyvb/network.py:3-96).

``fetch_network`` reproduces the reference's discovery order (it defines the Gauss-Seidel update
order); ``learn`` compiles that order into batched plate operations on the GPU.
"""
import numpy as np

from . import plate as _plate
from .nodes import (Addition, DiagonalGamma, Gamma, Gaussian, Multiplication, Wishart, hstack)


class Network(object):
    verbose = True

    def __init__(self, nodes=[], mode=None, algo=None):
        self.nodes = []
        self.mode, self.algo = mode, algo
        [self.addnode(n) for n in nodes]

    def addnode(self, n):
        """Add a node (or list of nodes) to the network"""
        if type(n) is list:
            self.nodes.extend(n)
        else:
            self.nodes.append(n)

    def find_iterable(self):
        """the nodes which are to be updated, in discovery order (network.py:35-37)"""
        self.iterable_nodes = [e for e in self.nodes if isinstance(e, (Gaussian, Gamma, DiagonalGamma, Wishart))]

    def fetch_network(self):
        """Find all nodes connected to the nodes in the network (network.py:58-96).

        Same breadth-first order as the reference (nodes are appended to the list being scanned);
        membership is tracked in an id-set instead of the reference's O(n^2) list scans."""
        n_start = len(self.nodes)
        seen = set(id(e) for e in self.nodes)

        def extend(cands):
            new = []
            for e in cands:
                if id(e) not in seen:
                    seen.add(id(e))
                    new.append(e)
            self.nodes.extend(new)

        i = 0
        while i < len(self.nodes):
            n = self.nodes[i]
            i += 1
            if isinstance(n, Gaussian):
                extend(n.children)
                extend([n.mean_parent, n.precision_parent])
            elif isinstance(n, (Addition, Multiplication)):
                extend(n.children)
                extend([n.A, n.B])
            elif isinstance(n, hstack):
                extend(n.children)
                extend(n.parents)
            if isinstance(n, (Gamma, DiagonalGamma, Wishart)):
                extend(n.children)
        if self.verbose:
            print("Found " + str(len(self.nodes) - n_start) + " new nodes.")

    def learn(self, niters, tol=1e-3):
        """Sweep the iterable nodes until the bound improves by less than tol (network.py:40-56);
        like the reference this also stops when the bound DEcreases."""
        self.find_iterable()
        if self.verbose:
            print('Found' + str(len(self.iterable_nodes)) + ' iterable nodes\n')
        if not self.iterable_nodes:
            return
        pl = _plate.bind(self.iterable_nodes[0], mode=self.mode, algo=self.algo)
        sched = pl.schedule(self.iterable_nodes)
        old_llb = -np.inf
        self.llb_trace = []
        for i in range(niters):
            pl.sweep(sched)
            self.llb = pl.elbo()
            self.llb_trace.append(self.llb)
            if self.verbose:
                print(niters - i, self.llb)
            if self.llb - old_llb < tol:
                if self.verbose:
                    print("Convergence!")
                break
            old_llb = self.llb
        pl.engine.check()
