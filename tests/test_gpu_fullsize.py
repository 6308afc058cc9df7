"""GPU: BASELINE.json's full single-GPU size (config 2: N=1M, D=256, q=16, 20 % missing) through
size-independent properties, plus oracle parity on a random sample of its rows."""
import numpy as np
import pytest
import torch

from helpers import numpy_stats, tensor_rel
from oracle.plate_oracle import PlateOracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["dmma", "auto"])
def big(request):
    """The N = 1M engine on the all-DMMA path and on the default path (auto = the INT8 tensor-core path, what bench.py times)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import bench
    from pyvb_b200 import PlateEngine
    N, D, q = 1000000, 256, 16
    dev = torch.device("cuda", 0)
    X = bench.make_data(torch, N, D, q, 0.2, 1234, dev)
    e = PlateEngine(X, q, mode="B", algo=request.param, keep_sigma=False, device=dev)
    assert e.use_i8 == (request.param == "auto") and e.use_i8_stats == (request.param == "auto")
    e.init_random(seed=4321)
    yield e
    assert e.i8_fallbacks() == (0, 0)          # normalised data: the accuracy guard never fires
    del e
    torch.cuda.empty_cache()


def test_fullsize_sweeps_and_determinism(big):
    e = big
    tr = [e.iterate() for _ in range(4)]
    assert np.all(np.isfinite(tr))
    e.check()
    # the Z step is a pure function of (X, Gw, tau): running it twice gives bit-identical rows
    e.update_Z()
    a = e.MZ.clone()
    e.update_Z()
    assert torch.equal(a, e.MZ)
    # statistics are deterministic too (fixed-order two-stage reduction)
    e._stats_fresh = False; e._ensure_stats(); s1 = e.stats.clone()
    e._stats_fresh = False; e._ensure_stats()
    assert torch.equal(s1, e.stats)


def test_fullsize_stats_additive_and_checksums(big):
    e = big
    e.update_Z()
    e._stats_fresh = False
    e._ensure_stats()
    full = e.stats.clone()
    L = e.L
    v = L.views(full.cpu().numpy())
    # checksum of checksums: column sums of the masked statistics against plain torch reductions
    obs = ~torch.isnan(e.X)
    assert abs(v["cnt"].sum() - float(obs.sum())) < 0.5
    assert tensor_rel(v["zsum"], e.Zbar.sum(0).cpu().numpy()) < 1e-10
    assert tensor_rel(v["S"], e.M2.sum(0).cpu().numpy()) < 1e-10
    x0 = torch.where(obs, e.X, torch.zeros((), dtype=e.X.dtype, device=e.X.device))
    assert tensor_rel(v["colx"], x0.sum(0).cpu().numpy()) < 1e-10
    # T1 summed over d == sum_n |O_n| <zz^T>_n
    w = obs.sum(1).to(torch.float64)
    assert tensor_rel(v["T1"].sum(0), (w[:, None] * e.M2).sum(0).cpu().numpy()) < 1e-10
    assert tensor_rel(v["Ast"].sum(0), (x0.sum(1)[:, None] * e.Zbar).sum(0).cpu().numpy()) < 1e-9
    # additivity over two row shards (what the all-reduce relies on), through the C-ABI on sub-ranges
    from pyvb_b200 import _cabi
    lib = e.lib
    acc = torch.zeros_like(full)
    cache = torch.zeros_like(e.xcache)
    for lo, hi in [(0, 400003), (400003, e.N)]:
        part = torch.zeros_like(full)
        rc = lib.pyvb_stats_f64(hi - lo, e.D, e.q, e.X.data_ptr() + lo * e.D * 8, e.D, 0, 0, 0,
                                e.Zbar.data_ptr() + lo * e.ldmz * 8, e.ldmz, e.M2.data_ptr() + lo * e.ldmz * 8, e.ldmz,
                                e.logdet.data_ptr() + lo * 8, part.data_ptr(), e.ws.data_ptr(), e.ws_bytes,
                                cache.data_ptr(), 0, 0, 0, None, e.algo, e._stream())
        _cabi.check(rc, "stats")
        acc += part
    fin = torch.isfinite(full)
    assert tensor_rel(acc[fin].cpu().numpy(), full[fin].cpu().numpy()) < 1e-11


def test_fullsize_rows_match_oracle_on_a_sample(big):
    e = big
    e.update_Z()
    rng = np.random.RandomState(0)
    idx = np.sort(rng.choice(e.N, 1500, replace=False))
    it = torch.as_tensor(idx, device=e.X.device)
    Xs = e.X[it].cpu().numpy()
    o = PlateOracle(Xs, e.q, mode="B")
    st = e.get_state_small()
    o.Wbar, o.Wvar, o.mu = st["Wbar"], st["Wvar"], st["mu"]
    o.qa, o.qb = st["tau"], 1.0
    o.update_Z()
    z = e.Zbar[it].cpu().numpy()
    assert tensor_rel(z, o.Zbar) < 1e-9
    m2 = e.M2[it].cpu().numpy()
    from oracle.plate_oracle import pack_sym
    assert tensor_rel(m2, pack_sym(o.M2())) < 1e-9
    ld = e.logdet[it].cpu().numpy()
    assert tensor_rel(0.5 / ld, o.qldZ) < 1e-9
    # and the statistics layout, on the sample, against the numpy definition
    ref = numpy_stats(Xs, o.Zbar, o.Sig, e.q)
    from pyvb_b200 import PlateEngine
    small = PlateEngine(Xs, e.q, mode="B", algo="dmma" if not e.use_i8 else "auto", device=e.X.device)
    small.set_state({"Zbar": o.Zbar, "Sig": o.Sig})
    small._ensure_stats()
    if small.use_i8_stats:                      # (the first call fills the X-only cache on the DMMA kernels)
        small._stats_fresh = False
        small._ensure_stats()
        assert small.i8_stats_calls == 1
    got = small.stats.cpu().numpy()
    L = small.L
    assert tensor_rel(got[:L.scal], ref[:L.scal]) < 1e-11


def test_fullsize_T1_matches_float64_matmul_on_column_blocks(big):
    """T1 / Bst / Ast of the full 1M rows (K3 of whichever path the fixture runs: its INT8 scales are global column maxima
    over all N rows) against torch float64 matmuls accumulated over row blocks."""
    e = big
    e.update_Z()
    e._stats_fresh = False
    e._ensure_stats()
    v = e.L.views(e.stats)
    T1 = torch.zeros(e.D, e.P, dtype=torch.float64, device=e.X.device)
    Bst = torch.zeros(e.D, e.q, dtype=torch.float64, device=e.X.device)
    Ast = torch.zeros_like(Bst)
    step = 1 << 17
    for lo in range(0, e.N, step):
        x = e.X[lo:lo + step]
        o = (~torch.isnan(x)).to(torch.float64)
        x0 = torch.nan_to_num(x, nan=0.0)
        T1 += o.t() @ e.M2[lo:lo + step]
        Bst += o.t() @ e.Zbar[lo:lo + step]
        Ast += x0.t() @ e.Zbar[lo:lo + step]
    for name, ref in (("T1", T1), ("Bst", Bst), ("Ast", Ast)):
        got = v[name]
        row = ((got - ref).abs().amax(1) / ref.abs().amax(1)).max().item()       # per data dimension
        assert row < 1e-10, (name, row)


# ---- the default path at the other two configurations' shapes: BASELINE.json config 3 (one GPU's share of its rows is
# 1.25M; 160k here) and config 4 (ARD, q = 64), five sweeps against the oracle on a row sample + properties at full N
@pytest.mark.parametrize("N,D,q,miss,ard", [(160000, 1024, 32, 0.3, False), (60000, 512, 64, 0.3, True)])
def test_c3_c4_shapes_default_path_sample_rows_match_oracle(N, D, q, miss, ard):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import bench
    from pyvb_b200 import PlateEngine
    from oracle.plate_oracle import pack_sym
    dev = torch.device("cuda", 0)
    X = bench.make_data(torch, N, D, q, miss, 77, dev)
    e = PlateEngine(X, q, mode="B", algo="auto", keep_sigma=False, device=dev, ard=ard)
    assert e.use_i8 and e.use_i8_stats
    e.init_random(seed=5)
    for _ in range(3):
        e.iterate()
    e.check()
    f0 = e.i8_fallbacks()           # (the first sweeps from the stand-in start, tau = 1e8, may trip the guard: that is its job)
    # (1) one Z step of the sampled rows against the oracle, from the engine's current globals
    rng = np.random.RandomState(1)
    idx = np.sort(rng.choice(N, 1200, replace=False))
    it = torch.as_tensor(idx, device=dev)
    Xs = e.X[it].cpu().numpy()
    st = e.get_state_small()
    o = PlateOracle(Xs, q, mode="B", ard=ard)
    o.Wbar, o.Wvar, o.mu = st["Wbar"], st["Wvar"], st["mu"]
    o.qa, o.qb = st["tau"], 1.0
    o.update_Z()
    e.update_Z()
    z, m2 = e.Zbar[it].cpu().numpy(), e.M2[it].cpu().numpy()
    ref_m2 = pack_sym(o.M2())
    assert tensor_rel(z, o.Zbar) < 1e-9 and tensor_rel(m2, ref_m2) < 1e-9
    rowerr = np.max(np.abs(m2 - ref_m2), axis=1) / np.max(np.abs(ref_m2), axis=1)
    assert rowerr.max() < 1e-9, rowerr.max()                                       # per row, not only tensor-wise
    assert tensor_rel(0.5 / e.logdet[it].cpu().numpy(), o.qldZ) < 1e-9
    # (2) the statistics of ALL rows against float64 matmuls, per data dimension
    e._stats_fresh = False
    e._ensure_stats()
    v = e.L.views(e.stats)
    o_ = (~torch.isnan(e.X)).to(torch.float64)
    T1 = o_.t() @ e.M2
    Ast = torch.nan_to_num(e.X, nan=0.0).t() @ e.Zbar
    assert ((v["T1"] - T1).abs().amax(1) / T1.abs().amax(1)).max().item() < 1e-10
    assert ((v["Ast"] - Ast).abs().amax(1) / Ast.abs().amax(1)).max().item() < 1e-10
    # (3) the W update from those statistics against its numpy definition
    from helpers import w_update_from_stats
    gl = e.get_state_small()
    Wref, Wvref = w_update_from_stats(e.stats.cpu().numpy(), D, q, gl["Wbar"], gl["mu"], gl["tau"], gl["alpha"])
    e.update_W()
    assert tensor_rel(e.Wbar.cpu().numpy(), Wref) < 1e-9 and tensor_rel(e.Wvar.cpu().numpy(), Wvref) < 1e-9
    assert e.i8_fallbacks() == f0 and f0[0] == 0 and f0[1] <= 1    # the checked passes ran on the INT8 kernels
