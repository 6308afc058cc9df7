"""BASELINE config 5: LDS VB smoother batched over 65,536 independent sequences (T=200, state dim 8, obs dim 5).
Prints one JSON line: sequences*iterations/s on the GPU (CUDA events), HBM roofline of the launch, and the CPU
baselines on the same box (literal reference on a few sequences, numpy restatement on a batch)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=65536)
ap.add_argument("--T", type=int, default=200)
ap.add_argument("--q", type=int, default=8)
ap.add_argument("--d", type=int, default=5)
ap.add_argument("--no-cpu", action="store_true")
a = ap.parse_args()

from pyvb_b200 import LDSEngine
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(5)
# synthetic observations generated on the device: a stable random system per sequence
A = torch.randn(a.B, a.q, a.q, generator=g, device=dev, dtype=torch.float64) * (0.6 / a.q ** 0.5)
C = torch.randn(a.B, a.d, a.q, generator=g, device=dev, dtype=torch.float64) * 3.0
x = torch.randn(a.B, a.q, generator=g, device=dev, dtype=torch.float64)
Y = torch.empty(a.B, a.T, a.d, device=dev, dtype=torch.float64)
for t in range(a.T):
    if t:
        x = torch.einsum("bki,bi->bk", A, x) + 0.2 * torch.randn(a.B, a.q, generator=g, device=dev, dtype=torch.float64)
    Y[:, t] = torch.einsum("bki,bi->bk", C, x) + 0.2 * torch.randn(a.B, a.d, generator=g, device=dev, dtype=torch.float64)
e = LDSEngine(Y, a.q, device=dev)
e.init_random(seed=1)
for _ in range(3):
    e.iterate()
torch.cuda.synchronize()


def timed(niters, reps):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); e.iterate(niters); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


ms1 = timed(1, 5)
ms10 = timed(10, 3) / 10.0
e.check()
peaks = {}
try:
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
except Exception:
    pass
hbm = float(peaks.get("hbm_gbs", 6650.0))
by = a.B * 8.0 * (a.T * a.d + 2 * a.T * a.q + 2 * (2 * a.q * a.q + 2 * a.d * a.q + 2 * a.q + 2 * a.d) + 3 * a.q * a.q)
line = {"metric": "LDS VB smoother sequences*iterations/sec", "workload": "B=%d independent sequences, T=%d, state dim %d, obs dim %d, FP64" % (a.B, a.T, a.q, a.d),
        "value_one_iteration_per_launch": a.B / (ms1 * 1e-3), "ms_per_iteration_one_per_launch": ms1,
        "value_ten_iterations_per_launch": a.B / (ms10 * 1e-3), "ms_per_iteration_ten_per_launch": ms10,
        "unit": "sequences*iterations/s",
        "roofline": {"bound": "hbm", "achieved": by / (ms1 * 1e-3) * 1e-9, "peak": hbm, "unit": "GB/s",
                     "frac": by / (ms1 * 1e-3) * 1e-9 / hbm, "bytes_per_launch": by,
                     "note": "one iteration per launch: the sequence, its states and parameters cross HBM once each way"}}
if not a.no_cpu:
    from oracle.make_ref import import_ref
    from oracle.lds_oracle import LDSOracle
    Yh = Y[:256].cpu().numpy()
    o = LDSOracle(Yh, a.q); o.iterate()
    t0 = time.perf_counter(); o.iterate(); o.iterate(); dt = (time.perf_counter() - t0) / 2
    line["cpu_baseline"] = {"kind": "port", "value": 256 / dt, "unit": "sequences*iterations/s", "cores": os.cpu_count(),
                            "sample": "numpy restatement, 256 sequences x 2 iterations"}
    pyvb = import_ref()
    if pyvb is not None:
        import io, contextlib
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
        from gen_golden_lds import build
        with contextlib.redirect_stdout(io.StringIO()):
            np.random.seed(0)
            m = build(pyvb, Yh[0], a.q)
            Xs = m["Xs"]
            def sweep():
                [x.update() for x in Xs]; Xs.reverse(); [x.update() for x in Xs]; Xs.reverse()
                [n.update() for n in m["As"]]; [n.update() for n in m["Cs"]]; m["Q"].update(); m["R"].update()
            sweep()
            t0 = time.perf_counter(); sweep(); dt = time.perf_counter() - t0
        line["cpu_baseline"]["reference"] = {"kind": "reference", "value": 1.0 / dt, "cores": 1,
                                             "sample": "literal pyvb, 1 sequence x 1 iteration in %.2f s" % dt}
print(json.dumps(line))
