"""Two-GPU test of the row-sharded path (needs 2 visible GPUs, otherwise skipped): one process per GPU, NCCL for
the plumbing, the statistics all-reduce done inside the second-stage kernel over NVLink peer memory
(pyvb_peers).  The sharded run must reproduce the oracle on the FULL data sweep by sweep, both ranks must hold
bit-identical replicated state, and the peer exchange must agree with the plain NCCL all-reduce."""
import os
import socket

import numpy as np
import pytest

from helpers import tensor_rel
from oracle.plate_oracle import PlateOracle, synth_pca

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, comm, shape, mode, out_dir):
    import torch
    import torch.distributed as dist
    from pyvb_b200 import PlateEngine
    from pyvb_b200.dist import shard_rows
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["PYVB_COMM"] = comm
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    N, D, q = shape
    X = synth_pca(N, D, q, 0.3, seed=3)
    rng = np.random.RandomState(1)
    init = {"Wbar": rng.randn(D, q), "Wvar": np.ones((D, q)), "mu": np.zeros(D), "muvar": np.ones(D),
            "Zbar": rng.randn(N, q), "Sig": np.tile(np.eye(q), (N, 1, 1)), "qb": 0.5}
    if mode == "A":
        init["Xhat"] = np.where(np.isnan(X), rng.randn(N, D), X)
        init["V"] = np.where(np.isnan(X), 1.3, 0.0)
    lo, hi = shard_rows(N, world, rank)
    e = PlateEngine(X[lo:hi], q, mode=mode, device=dev, distributed=True, row_offset=lo)
    assert (e.peers is not None) == (comm == "peer")
    loc = dict(init)
    for k in ("Zbar", "Sig", "Xhat", "V"):
        if k in loc:
            loc[k] = loc[k][lo:hi]
    e.set_state(loc)
    elbo = [e.iterate() for _ in range(4)]
    st = e.get_state_small()
    e.check()
    np.savez(os.path.join(out_dir, "r%d_%s.npz" % (rank, comm)), elbo=np.array(elbo), Wbar=st["Wbar"], mu=st["mu"],
             qb=st["qb"], Zbar=e.Zbar.cpu().numpy())
    e.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode,shape", [("B", (3001, 64, 16)), ("B", (1500, 24, 5)), ("A", (801, 16, 3))])
def test_two_gpu_sharded_sweeps(tmp_path, mode, shape):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    N, D, q = shape
    for comm in ("peer", "nccl"):
        mp.spawn(_worker, args=(world, _free_port(), comm, shape, mode, str(tmp_path)), nprocs=world, join=True)
    X = synth_pca(N, D, q, 0.3, seed=3)
    rng = np.random.RandomState(1)
    init = {"Wbar": rng.randn(D, q), "Wvar": np.ones((D, q)), "mu": np.zeros(D), "muvar": np.ones(D),
            "Zbar": rng.randn(N, q), "Sig": np.tile(np.eye(q), (N, 1, 1)), "qb": 0.5}
    if mode == "A":
        init["Xhat"] = np.where(np.isnan(X), rng.randn(N, D), X)
        init["V"] = np.where(np.isnan(X), 1.3, 0.0)
    o = PlateOracle(X, q, mode=mode)
    o.load_state(init)
    ref = np.array([o.iterate() for _ in range(4)])
    got = {(r, c): np.load(os.path.join(str(tmp_path), "r%d_%s.npz" % (r, c))) for r in range(world) for c in ("peer", "nccl")}
    for c in ("peer", "nccl"):
        a, b = got[(0, c)], got[(1, c)]
        assert np.array_equal(a["elbo"], b["elbo"]) and np.array_equal(a["Wbar"], b["Wbar"])   # replicas bit-identical
        assert np.max(np.abs(a["elbo"] - ref) / np.abs(ref)) < 1e-9
        assert tensor_rel(a["Wbar"], o.Wbar) < 1e-9 and tensor_rel(a["mu"], o.mu) < 1e-9
        assert abs(float(a["qb"]) - o.qb) <= 1e-9 * abs(o.qb)
        Z = np.concatenate([a["Zbar"], b["Zbar"]], 0)
        assert tensor_rel(Z, o.Zbar) < 1e-9
    # own exchange vs NCCL: same sums up to the order of the additions
    assert np.max(np.abs(got[(0, "peer")]["elbo"] - got[(0, "nccl")]["elbo"]) / np.abs(ref)) < 1e-12
