"""The one exchange of a sweep: all-reduce(SUM) of the packed statistics buffer over torch.distributed.
NCCL over NVLink on the GPU box; the same call runs over gloo in the CPU tests."""


def allreduce_stats(t):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def shard_rows(n_total, world, rank):
    """Contiguous row block [lo, hi) of `rank`: blocks differ by at most one row."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class PeerExchange(object):
    """NVLink peer-memory exchange buffers of the statistics all-reduce (one process per GPU, one box).

    Every rank allocates one buffer through the C-ABI (cudaMalloc), publishes its CUDA IPC handle with one
    all_gather over torch.distributed, and opens the other ranks' buffers.  pyvb_stats_f64 then does the
    all-reduce inside its own second-stage kernel (include/pyvb_b200.h: pyvb_peers); NCCL is only the plumbing
    that carries the 64-byte handles."""

    def __init__(self, lib, stats_len, device):
        import ctypes
        import torch
        import torch.distributed as dist
        from . import _cabi
        self.lib, self.device = lib, device
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        nbytes = int(lib.pyvb_peer_bytes(stats_len))
        own = ctypes.c_void_p()
        _cabi.check(lib.pyvb_peer_alloc(nbytes, ctypes.byref(own)), "pyvb_peer_alloc")
        self.own = own.value
        h = ctypes.create_string_buffer(64)
        _cabi.check(lib.pyvb_peer_export(self.own, h), "pyvb_peer_export")
        mine = torch.tensor(list(h.raw), dtype=torch.uint8, device=device)
        allh = [torch.zeros(64, dtype=torch.uint8, device=device) for _ in range(self.world)]
        dist.all_gather(allh, mine)
        ptrs, self.opened = [], []
        for r in range(self.world):
            if r == self.rank:
                ptrs.append(self.own)
                continue
            raw = bytes(allh[r].cpu().tolist())
            p = ctypes.c_void_p()
            _cabi.check(lib.pyvb_peer_import(raw, ctypes.byref(p)), "pyvb_peer_import")
            ptrs.append(p.value)
            self.opened.append(p.value)
        self.bufs = torch.tensor(ptrs, dtype=torch.int64, device=device)
        self.epoch = 0
        self.struct = _cabi.Peers()
        self.struct.bufs, self.struct.world, self.struct.rank = self.bufs.data_ptr(), self.world, self.rank
        dist.barrier()

    def next(self):
        """The pyvb_peers argument of the next exchange (same epoch sequence on every rank)."""
        import ctypes
        self.epoch += 1
        self.struct.epoch = self.epoch
        return ctypes.byref(self.struct)

    def close(self):
        import torch
        import torch.distributed as dist
        if self.own is None:
            return
        torch.cuda.synchronize(self.device)
        if dist.is_initialized():
            dist.barrier()
        for p in self.opened:
            self.lib.pyvb_peer_close(p)
        self.lib.pyvb_peer_free(self.own)
        self.own, self.opened = None, []
