"""The general dense-constraint kernel (the literal rtmg contract: f2 = f Ri dense, c x d) at the
cfg4 size (c = 2999, d = 2000, F = 48 MB), next to the box-form kernel on the same problem."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
import lineqgpr_b200 as lqg
from lineqgpr_b200 import workloads as W
from lineqgpr_b200.samplers import whiten_box, EPSILON
par, _ = W.cfg4()
mu, S, lb, ub = par["mu"], par["Sigma"], par["lb"], par["ub"]
d = mu.size
init = np.minimum(np.maximum(par["map"], lb + EPSILON), ub - EPSILON)
U = whiten_box(mu, S, "ul")
import scipy.linalg as sla
finl, finu = np.isfinite(lb), np.isfinite(ub)
F = np.vstack([U[finl], -U[finu]])                      # f2 = [I; -I][finite rows] Ri
g = np.concatenate([(mu - lb)[finl], (ub - mu)[finu]])  # g2 = f mu + g
z0 = sla.solve_triangular(U, init - mu, lower=False)    # initial2 = R (x0 - mu)
for mode, chains, n in ((1, 296, 8), (2, 296, 8), (2, 592, 8), (2, 2368, 8), (2, 2368, 40)):
    for rep in range(2):
        t = time.perf_counter(); z, st = lqg.rtmg_canonical(n, 5, z0, F, g, n_chains=chains, burn_in=2, max_restarts=10000, mode=mode); wall = time.perf_counter() - t
    assert (F @ z[:, :50] + g[:, None]).min() >= -1e-9
    print(json.dumps(dict(kernel="canonical cfg4 (dense F 2999 x 2000)", mode={1: "per-chain CTA", 2: "lock-step GEMM rounds"}[mode], chains=chains, n_per=n, chain_ms=round(st["chain_ms"], 1),
          bounces_per_tr=round(st["bounces"] / st["transitions"], 2), tr_per_s=round(st["transitions"] / (st["chain_ms"] * 1e-3)),
          viol=st["violations"], wall_s=round(wall, 2))), flush=True)
