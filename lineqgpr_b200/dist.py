"""Multi-GPU layer: one process per GPU, torch.distributed for the plumbing.

The path shards naturally: chains are independent given (mu, Sigma, lb, ub), so
rank r of W runs global chains [r*C/W, (r+1)*C/W) with Philox streams keyed by the GLOBAL chain id
(lqg_opts.chain_offset) — the union over ranks is bit-identical to the single-GPU job.  There is
no collective on the sampling path.  The one exchange step is the gather of summary bands
(posterior mean / quantiles): each rank summarises the test points it owns and the small band
arrays are all-gathered (NCCL over NVLink on GPUs; gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def shard_chains(n_chains: int, rank: int, world: int):
    """Contiguous chain range [lo, hi) of `rank`; remainders go to the low ranks."""
    base, rem = divmod(n_chains, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_rows(n_rows: int, rank: int, world: int):
    """Contiguous test-point range [lo, hi) owned by `rank` for the band summaries."""
    return shard_chains(n_rows, rank, world)


def simulate_rows(n_rows: int, rank: int, world: int):
    """Test-point range [lo, hi) that lqg_simulate summarises on `rank`: equal blocks of
    ceil(n_rows / world) rows (the band all-gather moves equal-sized pieces), the last ranks may own
    fewer rows or none."""
    per = -(-n_rows // world) if n_rows > 0 else 0
    lo = min(n_rows, rank * per)
    return lo, min(n_rows, lo + per)


class Comm:
    """lqg_comm of the C ABI: the NCCL communicator lqg_simulate uses for its two exchanges (xi shards,
    band pieces).  One process per GPU; the current CUDA device must be set before construction.
    The 128-byte NCCL id is created on rank 0 and handed over by `exchange`, a callable that
    broadcasts a bytes object from rank 0 (from_torch_distributed uses the default process group)."""

    def __init__(self, rank: int, world: int, exchange=None):
        from . import _lib as L
        self.rank, self.world = int(rank), int(world)
        idb = None
        if self.world > 1:
            buf = (L.C.c_char * 128)()
            if self.rank == 0:
                L.check(L.lib().lqg_comm_unique_id(buf), "lqg_comm_unique_id")
            idb = exchange(bytes(buf.raw))
            if len(idb) != 128:
                raise ValueError("the exchanged NCCL id must be 128 bytes")
        h = L.C.c_void_p()
        keep = L.C.create_string_buffer(idb, 128) if idb is not None else None
        L.check(L.lib().lqg_comm_init(keep, self.rank, self.world, L.C.byref(h)), "lqg_comm_init")
        self.handle = h

    @classmethod
    def from_torch_distributed(cls, group=None):
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)

        def exchange(b):
            box = [b]
            dist.broadcast_object_list(box, src=0, group=group)
            return box[0]
        return cls(rank, world, exchange)

    def close(self):
        from . import _lib as L
        if getattr(self, "handle", None) is not None and self.handle.value:
            L.lib().lqg_comm_destroy(self.handle)
            self.handle = None


def hmc_control_for_rank(control: dict, n_chains: int, rank: int, world: int) -> dict:
    """control list for this rank's tmvrnorm call: its share of the chains, global chain offset."""
    lo, hi = shard_chains(n_chains, rank, world)
    c = dict(control)
    c["n.chains"] = hi - lo
    c["chain.offset"] = int(control.get("chain.offset", 0)) + lo
    return c, (lo, hi)


def allgather_bands(mean_local, qtls_local, n_rows: int, group=None):
    """All-gathers per-rank band pieces (rows owned by each rank, in rank order) into the full
    (mean[n_rows], qtls[nprobs, n_rows]).  Works on CPU tensors (gloo) and CUDA tensors (NCCL)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    mean_local = torch.as_tensor(mean_local, dtype=torch.float64)
    qtls_local = torch.as_tensor(qtls_local, dtype=torch.float64)
    nprobs = qtls_local.shape[0]
    dev = mean_local.device
    sizes = [shard_rows(n_rows, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    # fixed-size pieces (all_gather needs equal shapes): [1 + nprobs, width], padded with NaN
    piece = torch.full((1 + nprobs, width), float("nan"), dtype=torch.float64, device=dev)
    lo, hi = sizes[rank]
    piece[0, : hi - lo] = mean_local
    piece[1:, : hi - lo] = qtls_local
    out = [torch.empty_like(piece) for _ in range(world)]
    dist.all_gather(out, piece, group=group)
    mean = torch.cat([out[r][0, : sizes[r][1] - sizes[r][0]] for r in range(world)])
    qtls = torch.cat([out[r][1:, : sizes[r][1] - sizes[r][0]] for r in range(world)], dim=1)
    return mean, qtls


def allreduce_moments(sum_local, sumsq_local, count_local, group=None):
    """Exact posterior mean / variance of eta from per-rank partial sums (all-reduce SUM)."""
    import torch
    import torch.distributed as dist
    buf = torch.cat([torch.as_tensor(sum_local, dtype=torch.float64).ravel(),
                     torch.as_tensor(sumsq_local, dtype=torch.float64).ravel(),
                     torch.tensor([float(count_local)], dtype=torch.float64, device=torch.as_tensor(sum_local).device)])
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    d = (buf.numel() - 1) // 2
    n = buf[-1]
    mean = buf[:d] / n
    var = (buf[d:2 * d] - n * mean * mean) / (n - 1)
    return mean, var
