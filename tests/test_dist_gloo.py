"""world_size-2 gloo test (CPU) of the multi-GPU path's host logic: rows shard, each rank builds the packed
statistics of its block, ONE all-reduce(SUM) through pyvb_b200.dist.allreduce_stats gives every rank the
full-data statistics, and the replicated W update computed from them equals the oracle's on the full data."""
import os
import socket

import numpy as np
import torch
import torch.multiprocessing as mp

from helpers import numpy_stats, tensor_rel, w_update_from_stats
from oracle.plate_oracle import PlateOracle, synth_pca


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, N, D, q, out_dir):
    import torch.distributed as dist
    from pyvb_b200.dist import allreduce_stats, shard_rows
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    X = synth_pca(N, D, q, 0.3, seed=5)
    rng = np.random.RandomState(7)
    Zbar = rng.randn(N, q)
    Sig = np.tile(np.eye(q), (N, 1, 1)) * (rng.rand(N, 1, 1) + 0.5)
    lo, hi = shard_rows(N, world, rank)
    t = torch.from_numpy(numpy_stats(X[lo:hi], Zbar[lo:hi], Sig[lo:hi], q))
    allreduce_stats(t)
    np.save(os.path.join(out_dir, "stats_%d.npy" % rank), t.numpy())
    dist.destroy_process_group()


def test_sharded_stats_allreduce_gloo(tmp_path):
    N, D, q, world = 301, 12, 3, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, N, D, q, str(tmp_path)), nprocs=world, join=True)
    X = synth_pca(N, D, q, 0.3, seed=5)
    rng = np.random.RandomState(7)
    Zbar = rng.randn(N, q)
    Sig = np.tile(np.eye(q), (N, 1, 1)) * (rng.rand(N, 1, 1) + 0.5)
    full = numpy_stats(X, Zbar, Sig, q)
    got = [np.load(os.path.join(str(tmp_path), "stats_%d.npy" % r)) for r in range(world)]
    assert np.array_equal(got[0], got[1])                       # every rank holds the same reduced buffer
    assert tensor_rel(got[0], full) < 1e-13
    # the replicated update driven by the reduced buffer == the oracle's W update on the full data
    o = PlateOracle(X, q, mode="B")
    o.Zbar, o.Sig = Zbar.copy(), Sig.copy()
    rng2 = np.random.RandomState(3)
    o.Wbar = rng2.randn(D, q)
    o.mu = rng2.randn(D) * 0.1
    W0, mu0, tau = o.Wbar.copy(), o.mu.copy(), o.tau
    o.update_W()
    W, Wvar = w_update_from_stats(got[0], D, q, W0, mu0, tau, o.alpha0)
    assert tensor_rel(W, o.Wbar) < 1e-12 and tensor_rel(Wvar, o.Wvar) < 1e-12


def test_shard_rows_partition():
    from pyvb_b200.dist import shard_rows
    for n, w in [(10, 3), (8, 8), (5, 8), (1000003, 8)]:
        blocks = [shard_rows(n, w, r) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= 1
