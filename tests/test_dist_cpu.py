"""CPU tests (gloo, world_size 2) of the multi-GPU host logic: chain sharding and the band gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_chains_partition():
    from lineqgpr_b200.dist import shard_chains
    for n, w in [(1000, 8), (7, 2), (3, 4), (1184, 8), (1, 1)]:
        parts = [shard_chains(n, r, w) for r in range(w)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(parts[r][1] == parts[r + 1][0] for r in range(w - 1))
        sizes = [hi - lo for lo, hi in parts]
        assert max(sizes) - min(sizes) <= 1


def test_control_for_rank_offsets():
    from lineqgpr_b200.dist import hmc_control_for_rank
    seen = []
    for r in range(4):
        c, (lo, hi) = hmc_control_for_rank({"burn.in": 5, "seed": 3}, 10, r, 4)
        assert c["chain.offset"] == lo and c["n.chains"] == hi - lo and c["seed"] == 3
        seen += list(range(lo, hi))
    assert seen == list(range(10))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, n_rows, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lineqgpr_b200.dist import allgather_bands, allreduce_moments, shard_rows
    from oracle import lineqgpr_oracle as O
    rng = np.random.default_rng(0)
    y = rng.standard_normal((n_rows, 501))            # every rank holds all samples (after the xi all-gather)
    lo, hi = shard_rows(n_rows, rank, world)
    m_loc, q_loc = O.bands(y[lo:hi], (0.025, 0.975))  # stand-in for lqg_bands on this rank's rows
    mean, q = allgather_bands(m_loc, q_loc, n_rows)
    m_all, q_all = O.bands(y, (0.025, 0.975))
    ok = np.array_equal(mean.numpy(), m_all) and np.array_equal(q.numpy(), q_all)
    # moments from sample shards: rank r holds columns r::world
    cols = y[:, rank::world]
    mean2, var2 = allreduce_moments(cols.sum(1), (cols ** 2).sum(1), cols.shape[1])
    ok = ok and np.allclose(mean2.numpy(), y.mean(1), rtol=1e-12, atol=1e-14) and np.allclose(var2.numpy(), y.var(1, ddof=1), rtol=1e-10)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize("n_rows", [7, 16])
def test_band_gather_world2_gloo(n_rows):
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, n_rows, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_simulate_rows_cover_all_test_points():
    """Row partition of lqg_simulate (equal blocks of ceil(n_t / world) rows: its band all-gather moves
    equal-sized pieces), mirrored in lineqgpr_b200.dist.simulate_rows."""
    from lineqgpr_b200.dist import simulate_rows
    for n, w in [(1000, 8), (100000, 8), (7, 2), (3, 4), (0, 2), (1, 1), (9, 8)]:
        parts = [simulate_rows(n, r, w) for r in range(w)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(parts[r][1] == parts[r + 1][0] for r in range(w - 1))
        per = -(-n // w) if n else 0
        assert all(hi - lo <= per for lo, hi in parts)


def test_comm_of_one_rank_needs_no_nccl():
    """world == 1: the communicator is a plain handle (no NCCL id, no device needed to create it)."""
    from lineqgpr_b200.dist import Comm
    from lineqgpr_b200 import _lib as L
    c = Comm(0, 1)
    assert L.lib().lqg_comm_world(c.handle) == 1 and L.lib().lqg_comm_rank(c.handle) == 0
    c.close()
    import ctypes
    h = ctypes.c_void_p()
    assert L.lib().lqg_comm_init(None, 2, 2, ctypes.byref(h)) == 1          # rank out of range
    assert L.lib().lqg_comm_init(None, 0, 2, ctypes.byref(h)) == 1          # two ranks need an id


def _worker_unit_schedule(rank, world, port, ret):
    """The all-gather schedule of lqg_simulate restated on CPU tensors: ranks finish their sampling
    rounds at different moments, but every rank issues one gather per UNIT of samples, in unit order —
    so the sequence of collectives is the same everywhere and the gathered column order is fixed."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_per, unit, n_chains, m = 70, 32, 3, 4
    rng = np.random.default_rng(100 + rank)
    xi_local = rng.standard_normal((m, n_chains * n_per))                   # chain-major columns
    # round boundaries differ per rank (data-dependent progress of the slowest chain)
    rounds = [0, 20 + 7 * rank, 45 + 3 * rank, n_per]
    gathered, next_unit = [], 0
    n_units = -(-n_per // unit)
    for upto in rounds[1:]:
        while next_unit < n_units:
            u0, u1 = next_unit * unit, min(n_per, (next_unit + 1) * unit)
            if u1 > upto:
                break
            cols = np.concatenate([np.arange(c * n_per + u0, c * n_per + u1) for c in range(n_chains)])
            piece = torch.from_numpy(np.ascontiguousarray(xi_local[:, cols]))
            out = [torch.empty_like(piece) for _ in range(world)]
            dist.all_gather(out, piece)
            gathered.append(torch.cat(out, dim=1).numpy())
            next_unit += 1
    xi_all = np.concatenate(gathered, axis=1)
    ret[rank] = (next_unit == n_units, xi_all.shape, float(xi_all.sum()))
    dist.destroy_process_group()


def test_unit_schedule_is_rank_independent_world2_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_unit_schedule, args=(world, port, ret), nprocs=world, join=True)
    r = dict(ret)
    assert r[0][0] and r[1][0] and r[0][1] == r[1][1] == (4, 2 * 3 * 70)
    assert r[0][2] == r[1][2]                                               # every rank holds the same gathered xi
