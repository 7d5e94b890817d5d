import sys, os, time
import numpy as np
sys.path.insert(0, ".")
os.environ["LQG_TRACE"] = "1"
import torch
from lineqgpr_b200 import _lib as L
import bench
P = bench.build_problem("cfg4")
d = P["d"]; dev = torch.device("cuda", 0)
f = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
Ut = torch.from_numpy(np.asfortranarray(P["U"]).T.copy()).to(dev); St = torch.from_numpy(np.asfortranarray(P["Sigma"]).T.copy()).to(dev)
mu_t, lb_t, ub_t, init_t = f(P["mu"]), f(P["lb"]), f(P["ub"]), f(P["init"])
n_chains, n_per, burn = 1000, 125, 100
out_t = torch.empty((n_chains * n_per, d), dtype=torch.float64, device=dev)
lib = L.lib()
for rep in range(3):
    st = L.HmcStats()
    o = L.make_opts(mem=L.MEM_DEVICE, stream=torch.cuda.current_stream().cuda_stream, seed=rep, n_chains=n_chains)
    t = time.perf_counter()
    rc = lib.lqg_hmc_box(d, mu_t.data_ptr(), Ut.data_ptr(), 1, St.data_ptr(), lb_t.data_ptr(), ub_t.data_ptr(), init_t.data_ptr(), 0, n_per, burn, o, out_t.data_ptr(), None, st)
    torch.cuda.synchronize()
    print("DEVICE rep", rep, "rc", rc, "wall ms", 1e3 * (time.perf_counter() - t), "chain", st.chain_ms, "pre", st.prepass_ms, file=sys.stderr)
