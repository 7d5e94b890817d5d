"""Throughput probe: RSM (cfg2), Phi.xi GEMM + bands at a large slab (secondary stages)."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
import lineqgpr_b200 as lqg
from lineqgpr_b200 import workloads as W
par, _ = W.cfg2(sampler="RSM")
for n in (10**4, 10**5):
    t = time.perf_counter(); x = lqg.tmvrnorm(par, n, {"seed": 1}); wall = time.perf_counter() - t
    s = dict(lqg.last_stats)
    print(json.dumps(dict(stage="RSM cfg2", nsim=n, wall_s=round(wall, 4), kernel_ms=round(s["kernel_ms"], 3), proposals=s["proposals"],
          in_box=s["in_box"], rounds=s["rounds"], proposals_per_s=round(s["proposals"] / (s["kernel_ms"] * 1e-3)),
          accepted_per_s=round(n / (s["kernel_ms"] * 1e-3)))), flush=True)
par, _ = W.cfg2(sampler="HMC")
for n in (10**4, 10**5):
    t = time.perf_counter(); x = lqg.tmvrnorm(par, n, {"seed": 1, "burn.in": 1}); wall = time.perf_counter() - t
    s = dict(lqg.last_stats)
    print(json.dumps(dict(stage="HMC cfg2", nsim=n, wall_s=round(wall, 4), chain_ms=round(s["chain_ms"], 3), n_chains=s["n_chains"],
          samples_per_s_kernel=round(n / (s["chain_ms"] * 1e-3)))), flush=True)
par, _ = W.cfg3()
t = time.perf_counter(); x = lqg.tmvrnorm(dict(par, **{"class": "HMC"}), 10**5, {"seed": 1, "burn.in": 1}); wall = time.perf_counter() - t
s = dict(lqg.last_stats)
print(json.dumps(dict(stage="HMC cfg3", nsim=10**5, wall_s=round(wall, 4), chain_ms=round(s["chain_ms"], 3), n_chains=s["n_chains"],
      samples_per_s_kernel=round(1e5 / (s["chain_ms"] * 1e-3)))), flush=True)
# paths: cfg5-shaped slab: n_t = 4000 test points, m = 1000, N = 2e5 samples
rng = np.random.default_rng(0)
_, aux = W.cfg5(n_test=4000)
xi = rng.standard_normal((1000, 200000))
t = time.perf_counter(); mean, q, info = lqg.paths_bands(aux["Phi_test"], xi, (0.025, 0.975), slab_rows=2000); wall = time.perf_counter() - t
print(json.dumps(dict(stage="paths+bands", n_t=4000, m=1000, N=200000, wall_s=round(wall, 3), **{k: (round(v, 3) if isinstance(v, float) else v) for k, v in info.items()})), flush=True)
