"""One lqg_hmc_box call on a headline shape, for ncu.  usage: profile_box.py [cfg4|cfg5] [chains] [n_per] [burn]"""
import sys
import numpy as np
sys.path.insert(0, ".")
import lineqgpr_b200 as lqg
from lineqgpr_b200 import workloads as W
from lineqgpr_b200.samplers import whiten_box
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
chains = int(sys.argv[2]) if len(sys.argv) > 2 else 592
n_per = int(sys.argv[3]) if len(sys.argv) > 3 else 16
burn = int(sys.argv[4]) if len(sys.argv) > 4 else 4
par = W.cfg4()[0] if cfg == "cfg4" else W.cfg5(n_test=10)[0]
U = whiten_box(np.asarray(par["mu"]), np.asarray(par["Sigma"]))
x = lqg.tmvrnorm(dict(par, **{"class": "HMC"}), chains * n_per, {"burn.in": burn, "n.chains": chains, "seed": 1, "U": U, "prepass": 1})
print({k: v for k, v in lqg.last_stats.items()})
