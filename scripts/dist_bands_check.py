"""2-rank NCCL check of the exchange step (run under torchrun on >= 2 GPUs):
every rank samples its share of the chains (lqg_hmc_box with chain_offset), xi shards are
all-gathered over NCCL, every rank summarises the test points it owns (lqg_dgemm + lqg_bands on
device memory), the band pieces are all-gathered.  Rank 0 compares with a single-process run."""
import os, sys, json
import numpy as np
sys.path.insert(0, ".")
import torch, torch.distributed as dist
import lineqgpr_b200 as lqg
from lineqgpr_b200 import workloads as W, dist as D
from lineqgpr_b200.simulate import lambda_pinv

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
par, aux = W.cfg5(D=20, knots=10, n=150, n_test=1001)
obj = dict(par, **{"class": "HMC"})
C, n_per = 64, 16
ctl, (lo, hi) = D.hmc_control_for_rank({"burn.in": 5, "seed": 42, "n.chains": C}, C, rank, world)
eta_loc = lqg.tmvrnorm(obj, (hi - lo) * n_per, ctl)                       # d x N_loc, this rank's chains
xi_loc = lqg.dgemm(lambda_pinv(aux["Lambda"]), eta_loc)                   # m x N_loc
# all-gather xi shards (equal sizes here) over NCCL
t_loc = torch.from_numpy(np.ascontiguousarray(xi_loc.T)).cuda()           # N_loc x m (row = sample)
parts = [torch.empty_like(t_loc) for _ in range(world)]
dist.all_gather(parts, t_loc)
xi_all = torch.cat(parts, 0).cpu().numpy().T                              # m x N
# bands for the rows this rank owns
n_t = aux["Phi_test"].shape[0]
r0, r1 = D.shard_rows(n_t, rank, world)
mean_loc, q_loc, info = lqg.paths_bands(aux["Phi_test"][r0:r1], xi_all, (0.025, 0.975))
mean, q = D.allgather_bands(torch.from_numpy(mean_loc).cuda(), torch.from_numpy(np.ascontiguousarray(q_loc)).cuda(), n_t)
if rank == 0:
    eta_one = lqg.tmvrnorm(obj, C * n_per, {"burn.in": 5, "seed": 42, "n.chains": C})
    xi_one = lqg.dgemm(lambda_pinv(aux["Lambda"]), eta_one)
    m1, q1, _ = lqg.paths_bands(aux["Phi_test"], xi_one, (0.025, 0.975))
    ok_xi = np.array_equal(xi_one, xi_all)
    ok_b = np.allclose(mean.cpu().numpy(), m1, rtol=1e-12, atol=1e-14) and np.array_equal(q.cpu().numpy(), q1)
    print(json.dumps({"world": world, "sharded_xi_equals_single_process": bool(ok_xi), "bands_equal": bool(ok_b),
                      "n_t": n_t, "N": C * n_per}))
    assert ok_xi and ok_b
dist.destroy_process_group()
