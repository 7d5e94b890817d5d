"""Wall time of the Python tmvrnorm wrapper (host numpy in/out) at the headline shape."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
import lineqgpr_b200 as lqg
from lineqgpr_b200 import workloads as W
par, _ = W.cfg4()
obj = dict(par, **{"class": "HMC"})
for rep in range(3):
    t = time.perf_counter(); x = lqg.tmvrnorm(obj, 125000, {"seed": rep}); wall = time.perf_counter() - t
    st = dict(lqg.last_stats)
    print(json.dumps(dict(stage="tmvrnorm(HMC) cfg4 via Python wrapper (whitening + sampling + copy out)", nsim=125000, wall_s=round(wall, 3),
          chain_ms=round(st["chain_ms"], 1), prepass_ms=round(st["prepass_ms"], 1), samples_per_s=round(125000 / wall), viol=st["violations"],
          inbox=bool((x >= par["lb"][:, None]).all() and (x <= par["ub"][:, None]).all()))), flush=True)
    del x
