"""Throughput of the general dense-constraint kernel (lqg_hmc_canonical) vs the box kernel on cfg1."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
import lineqgpr_b200 as lqg
from lineqgpr_b200 import workloads as W
from lineqgpr_b200.samplers import whiten_box, EPSILON
par, _ = W.cfg1()
mu, S, lb, ub = par["mu"], par["Sigma"], par["lb"], par["ub"]
init = np.minimum(np.maximum(par["map"], lb + EPSILON), ub - EPSILON)
U = whiten_box(mu, S, "tmg")
R = np.linalg.inv(U)                                   # R = chol(M)
fin = np.isfinite(lb)
F = np.vstack([np.eye(100)[fin]]) @ U                  # f2 = f Ri
g = (mu - lb)[fin]                                     # g2 = f mu + g
z0 = R @ (init - mu)
for chains, n in ((1, 1000), (1776, 100), (3552, 100)):
    for rep in range(2):
        t = time.perf_counter(); z, st = lqg.rtmg_canonical(n, 5, z0, F, g, n_chains=chains, burn_in=10); wall = time.perf_counter() - t
    print(json.dumps(dict(kernel="canonical cfg1", chains=chains, n_per=n, chain_ms=round(st["chain_ms"], 3), bounces_per_tr=round(st["bounces"] / st["transitions"], 2),
          exact_per_search=round(st["exact_evals"] / st["hit_searches"], 2), tr_per_s=round(st["transitions"] / (st["chain_ms"] * 1e-3)), viol=st["violations"])), flush=True)
for chains, n in ((1, 1000), (2368, 100), (4736, 100)):
    for rep in range(2):
        x = lqg.tmvrnorm(dict(par, **{"class": "HMC"}), chains * n, {"burn.in": 10, "n.chains": chains, "seed": 5, "U": U})
    st = dict(lqg.last_stats)
    print(json.dumps(dict(kernel="box cfg1", chains=chains, n_per=n, chain_ms=round(st["chain_ms"], 3), bounces_per_tr=round(st["bounces"] / st["transitions"], 2),
          tr_per_s=round(st["transitions"] / (st["chain_ms"] * 1e-3)), viol=st["violations"])), flush=True)
