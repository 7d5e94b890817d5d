import sys, os, time
import numpy as np
sys.path.insert(0, ".")
os.environ["LQG_TRACE"] = "1"
import torch
from lineqgpr_b200 import _lib as L
import bench
P = bench.build_problem("cfg4")
d = P["d"]
lib = L.lib()
pin = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
hU = torch.from_numpy(np.asfortranarray(P["U"]).T.copy()).pin_memory(); hS = torch.from_numpy(np.asfortranarray(P["Sigma"]).T.copy()).pin_memory()
hmu, hlb, hub, hinit = pin(P["mu"]), pin(P["lb"]), pin(P["ub"]), pin(P["init"])
n_chains = lib.lqg_hmc_box_slots(d); n_per = -(-125000 // n_chains); burn = 100
hout = torch.empty((n_chains * n_per, d), dtype=torch.float64).pin_memory()
for rep in range(3):
    st = L.HmcStats()
    o = L.make_opts(mem=L.MEM_HOST, stream=torch.cuda.current_stream().cuda_stream, seed=rep, n_chains=n_chains)
    t = time.perf_counter()
    rc = lib.lqg_hmc_box(d, hmu.data_ptr(), hU.data_ptr(), 1, hS.data_ptr(), hlb.data_ptr(), hub.data_ptr(), hinit.data_ptr(), 0, n_per, burn, o, hout.data_ptr(), None, st)
    torch.cuda.synchronize()
    print("HOST rep", rep, "rc", rc, "wall ms", 1e3 * (time.perf_counter() - t), "chain", st.chain_ms, "pre", st.prepass_ms, file=sys.stderr)
