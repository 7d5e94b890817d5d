"""Large-N sanity: 1.2e6 samples at cfg4 (d = 2000): 2.4e9 output elements (> 2^31), device resident."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
import torch
from lineqgpr_b200 import _lib as L
import bench
P = bench.build_problem("cfg4")
d = P["d"]; dev = torch.device("cuda", 0)
lib = L.lib()
f = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
Ut = torch.from_numpy(np.asfortranarray(P["U"]).T.copy()).to(dev); St = torch.from_numpy(np.asfortranarray(P["Sigma"]).T.copy()).to(dev)
mu_t, lb_t, ub_t, init_t = f(P["mu"]), f(P["lb"]), f(P["ub"]), f(P["init"])
n_chains = 2 * lib.lqg_hmc_box_slots(d); n_per = 2028; burn = 100
N = n_chains * n_per
out_t = torch.empty((N, d), dtype=torch.float64, device=dev)
st = L.HmcStats()
o = L.make_opts(mem=L.MEM_DEVICE, stream=torch.cuda.current_stream().cuda_stream, seed=3, n_chains=n_chains)
t = time.perf_counter()
rc = lib.lqg_hmc_box(d, mu_t.data_ptr(), Ut.data_ptr(), 1, St.data_ptr(), lb_t.data_ptr(), ub_t.data_ptr(), init_t.data_ptr(), 0, n_per, burn, o, out_t.data_ptr(), None, st)
torch.cuda.synchronize(); wall = time.perf_counter() - t
L.check(rc, "lqg_hmc_box")
lb, ub = lb_t[None, :], ub_t[None, :]
last = out_t[-4096:]                      # the tail of the matrix lives beyond 2^31 elements
ok_tail = bool(((last >= lb) & (last <= ub)).all().item()) and bool(torch.isfinite(last).all().item())
m = out_t.mean(0).cpu().numpy()
print(json.dumps(dict(stage="large-N cfg4 device resident", N=N, elements=N * d, wall_s=round(wall, 3), samples_per_s=round(N / wall),
      violations=st.violations, failed=st.failed_chains, transitions=st.transitions, tail_ok=ok_tail,
      mean_in_box=bool(((m >= P["lb"]) & (m <= P["ub"])).all()))))
