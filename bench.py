#!/usr/bin/env python
"""Benchmark of the VB-PCA (missing data) hot path -- BASELINE.json's metric: rows*iterations/s.

    python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA arm (one process per GPU)
    python bench.py --impl reference [--steps K] [--warmup W]      the reference's own CPU implementation

Workload (config.workload): BASELINE.json configs[1] -- N=1M rows, D=256, q=16, 20% entries missing,
FP64, masked ("mode B") VB-PCA, synthetic data generated on the device.  With --gpus N every rank owns
its own 1M-row shard (weak scaling); the only exchange per sweep is one NCCL all-reduce of the packed
statistics buffer.  A "step" is one full VB sweep (W columns, Z rows, Mu, Beta, ELBO) over the shard.
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "VB-PCA missing-data rows*iters/sec"
UNIT = "rows*iters/s"


def workload_name(a):
    return "VB-PCA missing data N=%d (per GPU) D=%d q=%d %d%% missing FP64 mode=%s%s" % (
        a.N, a.D, a.q, int(round(a.missing * 100)), a.mode, " ARD" if getattr(a, "ard", False) else "")


# ----------------------------------------------------------------------------- clocks
class ClockSampler(object):
    """Polls NVML (clocks, power, throttle reasons) every 20 ms from a thread; the samples that fall inside the timed
    region are reported.  It is created BEFORE the barrier that precedes the timed region: nvmlInit takes ~0.1 s, and
    when only rank 0 paid it after the barrier every other rank's first exchange waited for rank 0 inside ITS timed
    region (0.3-9 ms per sweep of apparent multi-GPU overhead, tools/comm_probe.sh)."""
    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
               "hw_power_brake": 0x80, "sync_boost": 0x10}

    def __init__(self, index, interval=0.02):
        self.rows, self.ok, self._stop, self.interval = [], False, False, float(interval)
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        while not self._stop:
            try:
                fields = os.environ.get("PYVB_SAMPLER_FIELDS", "crp")      # diagnosis: which query disturbs
                rs, clk, pw = 0, 0, 0.0
                if "r" in fields:
                    try:
                        rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    except Exception:
                        rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                if "c" in fields:
                    clk = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                if "p" in fields:
                    pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.rows.append((time.time(), clk, pw, rs))
            except Exception:
                pass
            time.sleep(self.interval)

    def stop(self, t0, t1):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "")]}
        self._stop = True
        self.t.join(timeout=1.0)
        rows = [r for r in self.rows if t0 <= r[0] <= t1] or self.rows[-3:]
        sm = sorted(r[1] for r in rows)
        mask = 0
        for r in rows:
            mask |= int(r[3])
        reasons = sorted(n for n, bit in self.REASONS.items() if mask & bit)
        return {"sm_mhz": float(sm[len(sm) // 2]) if sm else None, "sm_max_mhz": float(self.max), "reasons": reasons,
                "power_w_max": max(r[2] for r in rows) if rows else None, "samples": len(rows)}


# ----------------------------------------------------------------------------- CPU baselines
def _synth_host(N, D, q, missing, seed):
    import numpy as np
    rng = np.random.RandomState(seed)
    W = rng.randn(D, q); Z = rng.randn(N, q); mu = rng.randn(D)
    X = Z @ W.T + mu[None, :] + rng.randn(N, D) * np.sqrt(1.0 / 20.0)
    X[rng.rand(N, D) < missing] = np.nan
    return X


def time_literal_reference(D, q, missing, rows, steps, warmup):
    """The UNMODIFIED reference (py3-translated copy in oracle/_ref), built exactly like
    examples/PCA_missing_data.py:31-42, one Network.learn(1) sweep per step."""
    import numpy as np
    from oracle.make_ref import import_ref
    pyvb = import_ref()
    if pyvb is None:
        return None
    X = _synth_host(rows, D, q, missing, seed=99)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        np.random.seed(0)
        nodes = pyvb.nodes
        Ws = [nodes.Gaussian(D, np.zeros((D, 1)), np.eye(D) * 1e-3) for i in range(q)]
        W = nodes.hstack(Ws)
        Mu = nodes.Gaussian(D, np.zeros((D, 1)), np.eye(D) * 1e-3)
        Beta = nodes.Gamma(D, 1e-3, 1e-3)
        Zs = [nodes.Gaussian(q, np.zeros((q, 1)), np.eye(q)) for i in range(rows)]
        Xs = [nodes.Gaussian(D, W * z + Mu, Beta) for z in Zs]
        [xn.observe(xv.reshape(D, 1)) for xn, xv in zip(Xs, X)]
        net = pyvb.Network(); net.addnode(W); net.fetch_network()
        for _ in range(warmup):
            net.learn(1, tol=-np.inf)
        t0 = time.perf_counter()
        for _ in range(steps):
            net.learn(1, tol=-np.inf)
        dt = time.perf_counter() - t0
    return {"value": rows * steps / dt, "seconds": dt, "rows": rows, "steps": steps}


def time_port(D, q, missing, rows, steps, warmup):
    """The numpy restatement (oracle/plate_oracle.py, mode B), all host cores through BLAS."""
    import numpy as np
    from oracle.plate_oracle import PlateOracle
    X = _synth_host(rows, D, q, missing, seed=98)
    o = PlateOracle(X, q, mode="B")
    rng = np.random.RandomState(0)
    o.Wbar = rng.randn(D, q); o.Zbar = rng.randn(rows, q)
    for _ in range(warmup):
        o.iterate()
    t0 = time.perf_counter()
    for _ in range(steps):
        o.iterate()
    dt = time.perf_counter() - t0
    return {"value": rows * steps / dt, "seconds": dt, "rows": rows, "steps": steps}


def cpu_baseline(a, budget_s=25.0):
    cores = os.cpu_count() or 1
    lit = None
    try:
        lit = time_literal_reference(a.D, a.q, a.missing, rows=max(1, int(budget_s / 2 / 7.0)), steps=1, warmup=0)
    except Exception as e:  # pragma: no cover
        lit = None
        sys.stderr.write("literal reference failed: %r\n" % (e,))
    port = time_port(a.D, a.q, a.missing, rows=16384, steps=1, warmup=1)
    if lit is not None:
        out = {"value": lit["value"], "unit": UNIT, "cores": 1, "kind": "reference",
               "sample": "literal pyvb (py3-translated, single-threaded Python+numpy): %d rows x %d sweep of the same "
                         "D=%d q=%d %d%%-missing workload in %.1f s (its Z step allocates 8*q^2*D^2 bytes per row)"
                         % (lit["rows"], lit["steps"], a.D, a.q, int(a.missing * 100), lit["seconds"])}
    else:
        out = {"value": port["value"], "unit": UNIT, "cores": cores, "kind": "port", "sample": ""}
    out["port"] = {"value": port["value"], "cores": cores,
                   "sample": "numpy plate oracle (mode B, OpenBLAS on all cores): %d rows x %d sweep in %.1f s"
                             % (port["rows"], port["steps"], port["seconds"])}
    return out


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    total = max(1, a.steps + a.warmup)
    rows = max(1, min(8, int(200.0 / (total * 7.0))))
    res, kind, cores = None, "reference", 1
    try:
        res = time_literal_reference(a.D, a.q, a.missing, rows=rows, steps=a.steps, warmup=a.warmup)
    except Exception as e:  # pragma: no cover
        sys.stderr.write("literal reference failed: %r\n" % (e,))
    if res is None:
        kind, cores = "port", os.cpu_count() or 1
        res = time_port(a.D, a.q, a.missing, rows=16384, steps=a.steps, warmup=a.warmup)
    sample = "%d rows per step, %d steps (+%d warm-up) of the D=%d q=%d %d%%-missing workload, %.1f s" % (
        res["rows"], res["steps"], a.warmup, a.D, a.q, int(a.missing * 100), res["seconds"])
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * res["seconds"] / max(1, a.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a), "sample": sample},
            "cpu_baseline": {"value": res["value"], "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- our arm
def make_data(torch, N, D, q, missing, seed, dev):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    f64 = torch.float64
    Wt = torch.randn(D, q, generator=g, device=dev, dtype=f64)
    mu = torch.randn(D, generator=g, device=dev, dtype=f64)
    X = torch.empty(N, D, device=dev, dtype=f64)
    step = 1 << 16
    for lo in range(0, N, step):
        n = min(step, N - lo)
        Z = torch.randn(n, q, generator=g, device=dev, dtype=f64)
        x = Z @ Wt.t() + mu + torch.randn(n, D, generator=g, device=dev, dtype=f64) * (1.0 / 20.0) ** 0.5
        m = torch.rand(n, D, generator=g, device=dev) < missing
        x[m] = float("nan")
        X[lo:lo + n] = x
    return X


def bench_f32(torch, a, dev, eng64, _cabi, lib, time_calls):
    """The FP32 variant on the same data: sweeps/s and its three kernels (CUDA events)."""
    from pyvb_b200 import PlateEngine
    X = eng64.X
    e = PlateEngine(X, a.q, mode="B", keep_sigma=False, device=dev, precision="f32")
    e.init_random(seed=4321, rank=0)
    for _ in range(3):
        e.iterate_async()
    torch.cuda.synchronize(dev)
    K = max(3, min(a.steps, 10))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    slots = [e.iterate_async() for _ in range(K)]
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / K
    e.check()
    st = torch.cuda.current_stream(dev).cuda_stream
    e._ensure_gw()

    def k1():
        _cabi.check(lib.pyvb_zstep_k1_f32(e.N, e.D, e.q, e.planes.data_ptr(), e.GT.data_ptr(), e.WT.data_ptr(),
                                          e.P0.data_ptr(), e.h0.data_ptr(), e.gl.data_ptr(), e.MZ.data_ptr(), st), "k1")
    ms_k1 = time_calls(k1, 5)
    ms_z = time_calls(lambda: e.update_Z(), 5)

    def stats_call():
        e._stats_fresh = False
        e._ensure_stats()
    ms_s = time_calls(stats_call, 5)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    ncp = e.ldmz
    by_k1 = e.N * (3.0 * 2 * e.D + 4.0 * ncp)           # the three bf16 planes once + the FP32 row out
    by_k3 = e.N * (3.0 * 2 * e.D + 3.0 * 2 * ncp)       # the planes + the bf16 x 3 rows once
    out = {"value": e.N / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "elbo_last": float(e.trace[slots[-1]].item()),
           "dtype": "f32 rows, bf16x3 tcgen05 contraction with FP32 TMEM accumulation, FP64 batched solve",
           "tolerance": "one sweep from a shared state: 5e-5 tensor-wise on W, mu, Z, Sigma; 5e-3 on qb; see tests/test_gpu_f32.py",
           "kernels": {"k1_ms": ms_k1, "zstep_ms": ms_z, "k2_ms": ms_z - ms_k1, "stats_ms": ms_s},
           "roofline_k1": {"bound": "hbm", "achieved": by_k1 / (ms_k1 * 1e-3) * 1e-9, "peak": hbm, "unit": "GB/s",
                           "frac": by_k1 / (ms_k1 * 1e-3) * 1e-9 / hbm, "bytes_per_launch": by_k1,
                           "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650"},
           "roofline_k3": {"bound": "hbm", "achieved": by_k3 / (ms_s * 1e-3) * 1e-9, "peak": hbm, "unit": "GB/s",
                           "frac": by_k3 / (ms_s * 1e-3) * 1e-9 / hbm, "bytes_per_launch": by_k3}}
    del e
    return out


def run_ours(a):
    # exactly ONE line on stdout: libraries (NCCL banner ...) get stderr, the JSON line gets the real stdout
    sys.stdout.flush()
    real_out = os.dup(1)
    os.dup2(2, 1)
    import torch
    from pyvb_b200 import PlateEngine, _cabi
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    lib = _cabi.lib()

    X = make_data(torch, a.N, a.D, a.q, a.missing, 1234 + rank, dev)
    eng = PlateEngine(X, a.q, mode=a.mode, algo=a.algo, keep_sigma=False, distributed=(world > 1),
                      row_offset=rank * a.N, device=dev, ard=a.ard)
    del X
    eng.init_random(seed=4321, rank=rank)
    if os.environ.get("PYVB_NOCOMM"):          # diagnosis only: time the sweeps without the exchange
        eng.distributed = False

    def sync():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local) if (rank == 0 and not os.environ.get("PYVB_NOSAMPLER")) else None
    for _ in range(max(a.warmup, 3)):
        eng.iterate_async()
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    slots = [eng.iterate_async() for _ in range(a.steps)]
    e1.record()
    sync()
    t1 = time.time()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    clocks = sampler.stop(t0, t1) if sampler is not None else None
    eng.check()
    elbo = [float(v) for v in eng.trace[slots].cpu().tolist()]
    value = world * a.N * a.steps / (ms * 1e-3)

    # ---- end to end through the public API with HOST buffers: every step uploads the shard from pinned
    # host memory, runs one sweep and reads the bound back
    Xh = torch.empty(a.N, a.D, dtype=torch.float64, pin_memory=True)
    Xh.copy_(eng.X)
    res_h = torch.empty(1, dtype=torch.float64, pin_memory=True)
    e2e_steps = max(1, min(a.steps, 8))
    sync()
    e0.record()
    for _ in range(e2e_steps):
        slot = eng.iterate_from_host(Xh)         # chunked upload overlapped with the Z step
        res_h.copy_(eng.trace[slot:slot + 1], non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
    e1.record()
    sync()
    ms2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_val = world * a.N * e2e_steps / (float(ms2.item()) * 1e-3)
    del Xh

    # ---- per-kernel timing (CUDA events on the launching stream) + the FP64 tensor roofline of this box
    def time_calls(fn, n):
        fn(); torch.cuda.synchronize(dev)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(n):
            fn()
        a1.record(); torch.cuda.synchronize(dev)
        return a0.elapsed_time(a1) / n

    eng._ensure_gw()
    ms_z = time_calls(lambda: eng.update_Z(), 5)
    ms_k1 = ms_k2 = ms_k1_i8 = ms_k1_eta = None
    stream = torch.cuda.current_stream(dev).cuda_stream
    if eng.use_i8:
        def k1_phase(ph):
            _cabi.check(lib.pyvb_zstep_i8_f64(eng.N, eng.D, eng.q, eng.X.data_ptr(), eng.D, eng.mask8.data_ptr(),
                                              eng.Wbar.data_ptr(), eng.Wvar.data_ptr(), eng.Gw.data_ptr(), eng.ldg,
                                              eng.P0.data_ptr(), eng.h0.data_ptr(), eng.gl.data_ptr(), eng.MZ.data_ptr(),
                                              eng.ldmz, eng.GI.data_ptr(), eng.gscale.data_ptr(), 0,
                                              eng.logdet.data_ptr(), 0, ph, stream), "zstep_i8")
        ms_k1_i8 = time_calls(lambda: k1_phase(2), 5)
        ms_k1_eta = time_calls(lambda: k1_phase(3), 5)
        ms_k1 = time_calls(lambda: k1_phase(1), 5)
        def k2_only():
            _cabi.check(lib.pyvb_zsolve_f64(eng.N, eng.q, eng.MZ.data_ptr(), eng.ldmz, 0, eng.logdet.data_ptr(),
                                            eng.gl.data_ptr(), eng.zsums.data_ptr() if eng.zsums is not None else 0,
                                            stream), "zsolve")
        k1_phase(1); ms_k2 = time_calls(k2_only, 1)
        eng.update_Z()                           # restore a consistent state
    elif lib.pyvb_algo_supported(2, a.D, a.q):
        def k1_only():
            alg, eng.algo = eng.algo, 3          # PYVB_ALGO_DMMA_K1: contraction only
            eng.update_Z()
            eng.algo = alg
        ms_k1 = time_calls(k1_only, 5)
        def k2_only():                           # in place on rows left as [qprec | eta] by k1_only
            _cabi.check(lib.pyvb_zsolve_f64(eng.N, eng.q, eng.MZ.data_ptr(), eng.ldmz, 0, eng.logdet.data_ptr(),
                                            eng.gl.data_ptr(), eng.zsums.data_ptr() if eng.zsums is not None else 0,
                                            torch.cuda.current_stream(dev).cuda_stream), "zsolve")
        k1_only(); ms_k2 = time_calls(k2_only, 1)
        eng.update_Z()                           # restore a consistent state

    def stats_call():
        eng._stats_fresh = False
        dflag, eng.distributed = eng.distributed, False
        eng._ensure_stats()
        eng.distributed = dflag
    ms_s = time_calls(stats_call, 5)
    scratch = torch.empty(148 * 2 * 256, dtype=torch.float64, device=dev)
    iters = 20000
    ms_p = min(time_calls(lambda: _cabi.check(lib.pyvb_bench_dmma_f64(148 * 2, iters, scratch.data_ptr(),
               torch.cuda.current_stream(dev).cuda_stream), "bench_dmma"), 2) for _ in range(3))
    peak_tf = 148 * 2 * 8 * iters * 8 * 512.0 / (ms_p * 1e-3) * 1e-12
    D, q = a.D, a.q
    P = q * (q + 1) // 2
    fl_k1 = a.N * (2.0 * D * P + 2.0 * D * q)                            # K1 (SURVEY 8d terms)
    fl_z = fl_k1 + a.N * (q ** 3 + 2.0 * q * q)                          # + K2
    fl_s = a.N * (2.0 * D * P + 4.0 * D * q)                              # K3
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    i8_kernels = None
    if eng.use_i8:
        # INT8 path: the FP64 tensor pipe is no longer the bound; every kernel of the sweep is a stream over HBM.
        # Algorithmic bytes per launch: rows in / rows out, nothing counted twice.
        ncz = (P + q + 31) // 32 * 32
        by = {"zstep_i8 (INT8 mask contraction -> qprec)": (a.N * (1.0 * D + 8.0 * P), ms_k1_i8),
              "zstep_dmma_kernel<ETA> (eta = O.(X-mu) @ W on the FP64 tensor cores)": (a.N * (8.0 * D + 8.0 * q), ms_k1_eta),
              "zsolve (K2: batched q x q Cholesky / inverse / solve)": (a.N * (16.0 * (P + q) + 8.0), ms_k2),
              "statistics (digitize + INT8 T1/Bst + DMMA Ast + reduce)":
                  (a.N * (8.0 * (P + q) + 7.0 * ncz + 1.0 * D + 7.0 * ncz + 8.0 * D + 8.0 * q), ms_s)}
        i8_kernels = {k: {"ms": v[1], "algorithmic_GB": v[0] * 1e-9, "GBps": v[0] / (v[1] * 1e-3) * 1e-9,
                          "frac_of_hbm_peak": v[0] / (v[1] * 1e-3) * 1e-9 / hbm} for k, v in by.items()}
    if eng.use_i8:
        # dominant single kernel of the INT8-path sweep: the batched solve (K2), HBM bound
        kname = "zsolve (K2: batched q x q Cholesky / inverse / solve, in place on the MZ rows)"
        by_dom, ms_dom = a.N * (16.0 * (P + q) + 8.0), ms_k2
        fl_dom = None
    elif ms_k1 is not None:
        kname, fl_dom, ms_dom = "zstep_dmma_kernel (K1: mask @ vec(G) contraction on FP64 tensor cores)", fl_k1, ms_k1
    else:
        kname, fl_dom, ms_dom = "zstep (generic K1+K2)", fl_z, ms_z
    ach = fl_dom / (ms_dom * 1e-3) * 1e-12 if fl_dom is not None else None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tpath):
        try:
            # measured DRAM bytes per row of K1 (ncu --set full at N = 151,552 rows) scaled to this launch's rows
            tj = json.load(open(tpath))
            traffic = tj.get(("zsolve%d_dram_bytes_per_row" if eng.use_i8 else "zstep_dmma%d_dram_bytes_per_row") % a.q)
            traffic = traffic * a.N if traffic is not None else None
        except Exception:
            traffic = None
    if eng.use_i8:
        ach_b = by_dom / (ms_dom * 1e-3) * 1e-9
        roofline = {"bound": "hbm", "kernel": kname, "achieved": ach_b, "peak": hbm, "unit": "GB/s", "frac": ach_b / hbm,
                    "traffic": traffic, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                    "bytes_per_launch": by_dom, "ms_per_launch": ms_dom,
                    "fp64_tensor_peak_tflops": peak_tf,
                    "sweep_fp64_equivalent_tflops": a.N * (4.0 * D * P + 6.0 * D * q + q ** 3 + 2.0 * q * q) * a.steps
                    / (ms * 1e-3) * 1e-12,
                    "note": "the mask contractions run as exact integer GEMMs on the INT8 tensor cores: the sweep's FP64-"
                            "equivalent rate exceeds the FP64 tensor roof; every kernel is a stream over HBM (see kernels)"}
    else:
      roofline = {"bound": "tensor", "kernel": kname,
                "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": traffic,
                "peak_source": "FP64 DMMA.8x8x4 loop measured in this run (MEASURED_PEAKS.json has no FP64 figure; "
                               "BASELINE.md section 4 asks for it to be measured on the box)",
                "flops_per_launch": fl_dom, "ms_per_launch": ms_dom}
    kernels = {"zstep_ms": ms_z, "zstep_k1_ms": ms_k1, "zstep_k1_i8_ms": ms_k1_i8, "zstep_k1_eta_ms": ms_k1_eta,
               "zsolve_k2_ms": ms_k2, "i8_path": i8_kernels, "zstep_tflops": fl_z / (ms_z * 1e-3) * 1e-12,
               "stats_ms": ms_s, "stats_tflops": fl_s / (ms_s * 1e-3) * 1e-12,
               "sweep_algorithmic_tflops": world * a.N * (4.0 * D * P + 6.0 * D * q + q ** 3 + 2.0 * q * q)
               * a.steps / (ms * 1e-3) * 1e-12 / world}

    # ---- FP32 variant (tcgen05 / TMEM / TMA) on the same shard, N = 1 only: whole sweeps + its kernels
    f32v = None
    if world == 1 and a.mode == "B" and a.q in (16, 32) and a.D % 32 == 0 and not a.no_f32:
        try:
            f32v = bench_f32(torch, a, dev, eng, _cabi, lib, time_calls)
        except Exception as ex:  # pragma: no cover
            f32v = {"error": repr(ex)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
                "warmup": max(a.warmup, 3), "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(a), "algo": ("i8" if eng.use_i8 else a.algo), "rows_total": world * a.N,
                           "l2": "inputs larger than L2 (X shard %.2f GB per GPU)" % (a.N * a.D * 8 / 1e9),
                           "parallelism": "rows sharded over %d GPU(s), one all-reduce of %d doubles per sweep"
                                          % (world, eng.L.len)},
                # per sweep, all-DMMA path: wupdate, pack_gw, zstep K1, zsolve K2, stats GEMM, reduce(+exchange), global;
                # INT8 path: wupdate, pack_gw, pack_g_i8, zstep_i8, pack_weta, eta, zsolve, colmax_reduce, digitize, stats_i8,
                # stats_x, reduce(+exchange), global
                "clocks": clocks, "gpu_launches": (13 if eng.use_i8 else 7) * a.steps,
                "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": a.N * a.D * 8, "d2h_bytes_per_step": 8,
                        "steps": e2e_steps},
                "roofline": roofline, "kernels": kernels, "elbo_last": elbo[-1] if elbo else None}
        if f32v is not None:
            line["f32_variant"] = f32v
        if world == 1 and not a.no_cpu:
            line["cpu_baseline"] = cpu_baseline(a)
        sys.stdout.flush()
        os.write(real_out, (json.dumps(line) + "\n").encode())
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--N", type=int, default=1000000)
    ap.add_argument("--D", type=int, default=256)
    ap.add_argument("--q", type=int, default=16)
    ap.add_argument("--missing", type=float, default=0.2)
    ap.add_argument("--mode", default="B", choices=["A", "B"])
    ap.add_argument("--algo", default="auto", choices=["auto", "generic", "dmma", "i8"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-f32", action="store_true", help="skip the FP32-variant leg")
    ap.add_argument("--ard", action="store_true", help="ARD Gamma precisions per latent column (config 4)")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
