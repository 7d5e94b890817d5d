"""Throughput probe of lqg_hmc_box on the headline shapes (not the benchmark; sizing aid)."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
import lineqgpr_b200 as lqg
from lineqgpr_b200 import workloads as W
from lineqgpr_b200.samplers import whiten_box

def run(name, par, n_chains, n_per, burn, prepass):
    t = time.perf_counter()
    U = whiten_box(np.asarray(par["mu"]), np.asarray(par["Sigma"]))
    t_w = time.perf_counter() - t
    out = []
    for rep in range(2):
        t = time.perf_counter()
        x = lqg.tmvrnorm(dict(par, **{"class": "HMC"}), n_chains * n_per,
                         {"burn.in": burn, "n.chains": n_chains, "seed": 1 + rep, "U": U, "prepass": prepass})
        wall = time.perf_counter() - t
        st = dict(lqg.last_stats)
    tr = st["transitions"]
    print(json.dumps(dict(cfg=name, chains=n_chains, n_per=n_per, burn=burn, prepass=prepass, whiten_s=round(t_w, 2),
          wall_s=round(wall, 3), chain_ms=round(st["chain_ms"], 2), prepass_ms=round(st["prepass_ms"], 2),
          bounces_per_tr=round(st["bounces"] / tr, 2), exact_per_search=round(st["exact_evals"] / st["hit_searches"], 2),
          vel=st["vel_restarts"], chk=st["chk_restarts"], viol=st["violations"],
          tr_per_s_kernel=round(tr / ((st["chain_ms"] + st["prepass_ms"]) * 1e-3)),
          samples_per_s_wall=round(n_chains * n_per / wall))), flush=True)

par4, _ = W.cfg4()
par5, _ = W.cfg5(n_test=10)
for prepass in (0, 1):
    run("cfg4", par4, 1184, 8, 2, prepass)
    run("cfg5", par5, 1184, 8, 2, prepass)
run("cfg4", par4, 2368, 40, 10, 1)
run("cfg5", par5, 2368, 40, 10, 1)
run("cfg4", par4, 592, 160, 40, 1)
