"""Pageable (numpy / R-like) result matrix: staged copy on / off, and bit-equality with a pinned run."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, ".")
import lineqgpr_b200 as lqg
from lineqgpr_b200 import workloads as W, _lib as L
par, _ = W.cfg4()
obj = dict(par, **{"class": "HMC"})
nsim = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
ref = None
for rep in range(3):
    for ctl in ({"seed": 7}, {"seed": 7, "whiten": "device"}):
        t = time.perf_counter(); x = lqg.tmvrnorm(obj, nsim, dict(ctl)); wall = time.perf_counter() - t
        st = dict(lqg.last_stats)
        if "whiten" not in ctl:
            if ref is None: ref = x.copy()
            same = bool(np.array_equal(ref, x))
        else:
            same = None
        print(json.dumps(dict(staged=os.environ.get("LQG_STAGED", "1"), ctl=ctl, nsim=nsim, wall_s=round(wall, 3), chain_ms=round(st["chain_ms"], 1),
                              prepass_ms=round(st["prepass_ms"], 1), viol=st["violations"], same_as_first=same)), flush=True)
        del x
# pinned output through the same wrapper: must be bit-identical to the pageable result
slots = L.lib().lqg_hmc_box_slots(par["mu"].size)
n_per = -(-nsim // slots)
out = L.empty_result(par["mu"].size, slots * n_per)
t = time.perf_counter(); x = lqg.tmvrnorm(obj, nsim, {"seed": 7, "out": out}); wall = time.perf_counter() - t
print(json.dumps(dict(pinned_out=True, wall_s=round(wall, 3), same_as_pageable=bool(np.array_equal(ref, x)))), flush=True)
