"""End-to-end simulate() at the cfg4 size (m = 1000 knots, d = 2000) with the set-up stage on the
host (numpy/LAPACK, as R does it) and on the device."""
import sys
import time
import numpy as np
sys.path.insert(0, ".")
import lineqgpr_b200 as lqg
from lineqgpr_b200 import workloads as W

nsim = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
par, aux = W.cfg4()
model, xtest = aux["model"], aux["xtest"]
model = dict(model); model["localParam"] = dict(model["localParam"], sampler="HMC")
lqg.simulate(model, nsim=300, seed=1, xtest=xtest, control={"setup": "device"})      # warm-up: pool, module load
for setup in ("host", "device", "host", "device"):
    t0 = time.perf_counter()
    sim = lqg.simulate(model, nsim=nsim, seed=1, xtest=xtest, control={"setup": setup})
    wall = time.perf_counter() - t0
    y = sim["ysim"]
    print(f"setup={setup:6s}: simulate wall {wall:.3f} s  (predict {sim['predtime']:.3f} s, sampler call {sim['simtime']:.3f} s), "
          f"violations {sim['stats']['violations']}, ysim {y.shape}, mean(ysim) {y.mean():.6f}, sd {y.std():.6f}", flush=True)
