"""Batched VB smoother for linear dynamic systems (BASELINE config 5) -- host side.

Mirrors, for B independent sequences at once, the manual sweep of the reference's LDS example
(/root/reference/examples/Linear_Dynamic_System.py:47-76): columns of A and C as Gaussians under hstack,
DiagonalGamma Q and R, X_0 ~ N(0, I), X_t ~ N(A X_{t-1}, Q), Y_t ~ N(C X_t, R).  One `iterate()` = all X_t forwards,
all X_t backwards, A columns, C columns, Q, R -- one kernel launch through the C-ABI (pyvb_lds_iterate_f64).
There is no CPU fallback.
"""
import numpy as np
import torch

from . import _cabi


class LDSEngine(object):
    KEYS = ("A", "Avar", "C", "Cvar", "Qa", "Qb", "Ra", "Rb", "X")

    def __init__(self, Y, q, alpha0=1e-3, a0=1e-3, b0=1e-3, device=None, A_known=None):
        if not torch.cuda.is_available():
            raise RuntimeError("pyvb_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = _cabi.lib()
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        f64, dev = torch.float64, self.device
        Yt = torch.as_tensor(Y)
        if Yt.dim() == 2:
            Yt = Yt[None]
        self.Y = Yt.to(device=dev, dtype=f64).contiguous()
        self.B, self.T, self.d = (int(v) for v in self.Y.shape)
        self.q = int(q)
        assert 1 <= self.q <= 8 and 1 <= self.d <= 8, "state and observation dimensions up to 8"
        assert 3 <= self.T <= int(self.lib.pyvb_lds_max_len()), "sequence length 3 .. %d" % int(self.lib.pyvb_lds_max_len())
        self.alpha0, self.a0, self.b0 = float(alpha0), float(a0), float(b0)
        B, T, d, q = self.B, self.T, self.d, self.q
        # deterministic stand-in initialisation; parity runs inject the reference's random state with set_state()
        self.A = torch.zeros(B, q, q, dtype=f64, device=dev)
        self.Avar = torch.ones(B, q, q, dtype=f64, device=dev)
        self.C = torch.zeros(B, d, q, dtype=f64, device=dev)
        self.Cvar = torch.ones(B, d, q, dtype=f64, device=dev)
        self.Qa = torch.full((B, q), self.a0 + 0.5 * (T - 1), dtype=f64, device=dev)
        self.Qb = torch.ones(B, q, dtype=f64, device=dev)
        self.Ra = torch.full((B, d), self.a0 + 0.5 * T, dtype=f64, device=dev)
        self.Rb = torch.ones(B, d, dtype=f64, device=dev)
        self.X = torch.zeros(B, T, q, dtype=f64, device=dev)
        self.Xcov3 = torch.zeros(B, 3, q, q, dtype=f64, device=dev)
        self.status = torch.zeros(1, dtype=f64, device=dev)
        # known entries of A (examples/LDS_knowns_in_A.py:72-74), NaN = free; (q, q) is shared by all sequences
        self.A_known = None
        if A_known is not None:
            ak = torch.as_tensor(np.asarray(A_known, dtype=np.float64))
            if ak.dim() == 2:
                ak = ak[None].expand(B, q, q)
            assert tuple(ak.shape) == (B, q, q)
            self.A_known = ak.to(device=dev, dtype=f64).contiguous()

    def init_random(self, seed=0):
        g = torch.Generator(device=self.device)
        g.manual_seed(int(seed))
        self.A.normal_(generator=g).mul_(0.3)
        self.C.normal_(generator=g)
        self.X.normal_(generator=g)
        self.Avar.fill_(1.0)
        self.Cvar.fill_(1.0)
        self.Qb.fill_(0.5)
        self.Rb.fill_(0.5)

    def set_state(self, st):
        for k in self.KEYS:
            if k in st:
                dst = getattr(self, k)
                dst.copy_(torch.as_tensor(np.ascontiguousarray(st[k]), dtype=torch.float64).reshape(dst.shape).to(self.device))

    def get_state(self):
        out = {k: getattr(self, k).cpu().numpy() for k in self.KEYS}
        c3 = self.Xcov3.cpu().numpy()
        cov = np.repeat(c3[:, 1:2], self.T, axis=1)
        cov[:, 0] = c3[:, 0]
        cov[:, -1] = c3[:, 2]
        out["Xcov"] = cov
        return out

    def iterate(self, niters=1):
        """`niters` iterations of the reference's sweep (Linear_Dynamic_System.py:69-76) for every sequence."""
        p = lambda t: 0 if t is None else t.data_ptr()
        with torch.cuda.device(self.device):
            rc = self.lib.pyvb_lds_iterate_known_f64(self.B, self.T, self.q, self.d, p(self.Y), p(self.X), p(self.Xcov3),
                                                     p(self.A), p(self.Avar), p(self.C), p(self.Cvar), p(self.Qa), p(self.Qb),
                                                     p(self.Ra), p(self.Rb), p(self.A_known), self.alpha0, self.a0, self.b0,
                                                     int(niters), p(self.status),
                                                     torch.cuda.current_stream(self.device).cuda_stream)
        _cabi.check(rc, "pyvb_lds_iterate_known_f64")

    def check(self):
        n = float(self.status.item())
        if n > 0:
            raise np.linalg.LinAlgError("posterior precision of %d sequence(s) is not positive definite" % int(n))
