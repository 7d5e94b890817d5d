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
