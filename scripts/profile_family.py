"""One representative call per kernel family, for ncu (one family per invocation keeps the capture short).
usage: python scripts/profile_family.py rsm|expt|canon_tma|canon_batch|bands|gemm|simulate"""
import ctypes as C
import json
import sys
import numpy as np
sys.path.insert(0, ".")
import lineqgpr_b200 as lqg
from lineqgpr_b200 import _lib as L, workloads as W
from lineqgpr_b200.samplers import whiten_box, EPSILON

fam = sys.argv[1]
out = {"family": fam}
if fam == "rsm":                                   # cfg2: d = 20, ~7.6 % acceptance
    par, _ = W.cfg2(sampler="RSM")
    for _ in range(2):
        x = lqg.tmvrnorm(par, 200000, {"seed": 3})
    out.update({k: v for k, v in lqg.last_stats.items()})
elif fam == "expt":                                # cfg1: d = 100
    par, _ = W.cfg1(sampler="ExpT")
    for _ in range(2):
        x = lqg.tmvrnorm(par, 200000, {"seed": 3})
    out.update({k: v for k, v in lqg.last_stats.items()})
elif fam in ("canon_tma", "canon_batch"):
    par, _ = (W.cfg1() if fam == "canon_tma" else W.cfg4())
    mu, S, lb, ub = (np.asarray(par[k], float) for k in ("mu", "Sigma", "lb", "ub"))
    init = np.minimum(np.maximum(np.asarray(par["map"], float), lb + EPSILON), ub - EPSILON)
    U = whiten_box(mu, S, "ul")
    import scipy.linalg as sla
    z0 = sla.solve_triangular(U, init - mu, lower=False)
    d = mu.size
    rows = [np.eye(d)[np.isfinite(lb)], -np.eye(d)[np.isfinite(ub)]]
    F = np.vstack(rows) @ U                         # f2 = f Ri: dense c x d
    g = np.concatenate([(mu - lb)[np.isfinite(lb)], (ub - mu)[np.isfinite(ub)]])
    chains, n, mode = (1776, 50, 1) if fam == "canon_tma" else (296, 6, 2)
    for _ in range(2):
        z, st = lqg.rtmg_canonical(n, 5, z0, F, g, n_chains=chains, burn_in=2, mode=mode)
    out.update(st, c=int(F.shape[0]), d=int(d), chains=chains)
elif fam in ("bands", "gemm", "simulate"):
    from lineqgpr_b200.simulate import simulate_pipeline
    par, aux = W.cfg4(n_test=1000)
    for _ in range(2):
        r = simulate_pipeline(par, 296 * 128, Lambda=aux["Lambda"], Phi_test=aux["Phi_test"],
                              control={"burn.in": 4, "n.chains": 296, "seed": 1, "burn.in.exact": True})
    out.update({k: v for k, v in r["stats"].items()})
print(json.dumps(out))
