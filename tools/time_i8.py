"""Time the Z-step contraction (K1) of the all-DMMA path against the INT8 mask contraction + DMMA eta kernel, and a
whole sweep of both, on synthetic shards.  usage: python tools/time_i8.py [N D q missing]..."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from pyvb_b200 import PlateEngine, _cabi  # noqa: E402


def time_calls(fn, n):
    fn(); torch.cuda.synchronize()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for _ in range(n):
        fn()
    a1.record(); torch.cuda.synchronize()
    return a0.elapsed_time(a1) / n


def synth(N, D, q, missing, dev, seed=0):
    g = torch.Generator(device=dev); g.manual_seed(seed)
    W = torch.randn(D, q, generator=g, device=dev, dtype=torch.float64)
    X = torch.empty(N, D, device=dev, dtype=torch.float64)
    step = 1 << 17
    for lo in range(0, N, step):
        n = min(step, N - lo)
        Z = torch.randn(n, q, generator=g, device=dev, dtype=torch.float64)
        x = Z @ W.T + 0.1 * torch.randn(n, D, generator=g, device=dev, dtype=torch.float64)
        x[torch.rand(n, D, generator=g, device=dev) < missing] = float("nan")
        X[lo:lo + n] = x
    return X


def run(N, D, q, missing):
    dev = torch.device("cuda:0")
    X = synth(N, D, q, missing, dev)
    out = {"N": N, "D": D, "q": q, "missing": missing}
    res = {}
    for algo in ALGOS:
        e = PlateEngine(X, q, mode="B", algo=algo, keep_sigma=False)
        e.init_random(seed=5)
        for _ in range(2):
            e.iterate_async()
        torch.cuda.synchronize()
        e._ensure_gw()
        lib = e.lib
        if algo == "dmma":
            def k1():
                alg, e.algo = e.algo, 3
                e.update_Z()
                e.algo = alg
        else:
            def k1():
                _cabi.check(lib.pyvb_zstep_i8_f64(N, D, q, e.X.data_ptr(), D, e.mask8.data_ptr(), e.Wbar.data_ptr(),
                                                  e.Wvar.data_ptr(), e.Gw.data_ptr(), e.ldg, e.P0.data_ptr(),
                                                  e.h0.data_ptr(), e.gl.data_ptr(), e.MZ.data_ptr(), e.ldmz,
                                                  e.GI.data_ptr(), e.gscale.data_ptr(), 0, e.logdet.data_ptr(), 0, 1,
                                                  e._stream()), "zstep_i8")
        out["k1_%s_ms" % algo] = time_calls(k1, 5)
        res[algo] = e.MZ[: min(N, 200000)].clone()
        e.update_Z()
        out["zstep_%s_ms" % algo] = time_calls(lambda: e.update_Z(), 5)
        out["sweep_%s_ms" % algo] = time_calls(lambda: e.iterate_async(), 5)
        e.check()
        del e
        torch.cuda.empty_cache()
    if len(res) == 2:
        out["k1_rows_rel_diff"] = (res["dmma"] - res["i8"]).abs().max().item() / res["dmma"].abs().max().item()
    print(json.dumps(out), flush=True)


ALGOS = ("dmma", "i8")

if __name__ == "__main__":
    args = sys.argv[1:]
    if args and args[0] == "--only":
        ALGOS = (args[1],)
        args = args[2:]
    shapes = []
    while args:
        shapes.append((int(args[0]), int(args[1]), int(args[2]), float(args[3])))
        args = args[4:]
    for s in shapes or [(1000000, 256, 16, 0.2)]:
        run(*s)
