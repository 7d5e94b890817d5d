"""Config 5 at full size on one GPU, everything device resident:
10^6 constrained samples of the additive model (D = 100 x 10 knots, d = m = 1000), xi = Lambda^+ eta,
and mean / 2.5 % / 97.5 % bands of ysim = PhiAll.test %*% xi.sim at 10^5 test points — a
10^5 x 10^6 matrix (800 GB) that is never materialised: row slabs of the hat-basis matrix are built
on the device, multiplied on the FP64 tensor cores, summarised exactly and dropped.
usage: python scripts/run_cfg5_full.py [nsim] [n_test]"""
import ctypes as C
import json
import sys
import time
import numpy as np
sys.path.insert(0, ".")
import torch
import importlib
import lineqgpr_b200 as lqg
from lineqgpr_b200 import _lib as L, workloads as W, setup_stage as S
from lineqgpr_b200.samplers import EPSILON
sim_mod = importlib.import_module("lineqgpr_b200.simulate")

nsim = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10 ** 6
n_test = int(float(sys.argv[2])) if len(sys.argv) > 2 else 10 ** 5
t_all = time.perf_counter()
t0 = time.perf_counter()
par, aux = W.cfg5(n_test=n_test)
t_model = time.perf_counter() - t0
d = par["mu"].size
dev = torch.device("cuda", 0)
lib = L.lib()
stream = torch.cuda.current_stream().cuda_stream
o = lambda **kw: L.make_opts(mem=L.MEM_DEVICE, stream=stream, **kw)
up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
upF = lambda a: torch.from_numpy(np.asfortranarray(a).T.copy()).to(dev)          # column-major on the device

# ---- sampling: eta (d x nsim) stays on the device ----
slots = lib.lqg_hmc_box_slots(d)
n_per = -(-nsim // slots)
N = slots * n_per
mu, lb, ub = par["mu"], par["lb"], par["ub"]
init = np.minimum(np.maximum(par.get("map", mu), lb + EPSILON), ub - EPSILON)
mu_d, lb_d, ub_d, init_d, Sig_d = up(mu), up(lb), up(ub), up(init), upF(par["Sigma"])
eta_d = torch.empty((N, d), dtype=torch.float64, device=dev)                       # column-major d x N
st = L.HmcStats()
torch.cuda.synchronize(); t0 = time.perf_counter()
L.check(lib.lqg_hmc_box(d, mu_d.data_ptr(), None, 1, Sig_d.data_ptr(), lb_d.data_ptr(), ub_d.data_ptr(), init_d.data_ptr(), 0,
                        n_per, 100, o(seed=1, n_chains=slots), eta_d.data_ptr(), None, st), "hmc_box")
torch.cuda.synchronize(); t_sample = time.perf_counter() - t0

# ---- xi = Lambda^+ eta on the device ----
Lam = aux["Lambda"]
m = Lam.shape[1]
Lp_d = upF(S.lambda_pinv_device(Lam))                                               # m x d
xi_d = torch.empty((N, m), dtype=torch.float64, device=dev)
ms = C.c_double(0)
torch.cuda.synchronize(); t0 = time.perf_counter()
L.check(lib.lqg_dgemm(m, N, d, Lp_d.data_ptr(), m, eta_d.data_ptr(), d, xi_d.data_ptr(), m, None, o(), C.byref(ms)), "dgemm")
torch.cuda.synchronize(); t_xi = time.perf_counter() - t0
del eta_d

# ---- bands of PhiAll.test %*% xi.sim, slab by slab ----
ulist = aux["lineq_model"]["ulist"]
t0 = time.perf_counter()
mean, q, info = sim_mod.paths_bands(None, xi_d, (0.025, 0.975), xtest=aux["xtest"], ulist=ulist)
torch.cuda.synchronize(); t_bands = time.perf_counter() - t0

# ---- spot check of 3 test points against torch (sort-based) on the device ----
from lineqgpr_b200 import model as M
rows = [0, n_test // 2, n_test - 1]
Phi_rows = np.hstack([M.basisCompute_lineqGP(aux["xtest"][rows, k], ulist[k]) for k in range(len(ulist))])
y_rows = up(Phi_rows) @ xi_d.T                                                        # 3 x N
ref_mean = y_rows.mean(dim=1).cpu().numpy()
srt = torch.sort(y_rows, dim=1).values
def q7(p):
    h = (N - 1) * p; lo = int(np.floor(h)); w = h - lo
    return ((1 - w) * srt[:, lo] + w * srt[:, min(lo + 1, N - 1)]).cpu().numpy()
ok = bool(np.allclose(mean[rows], ref_mean, rtol=1e-10) and np.allclose(q[0, rows], q7(0.025), rtol=1e-9, atol=1e-12)
          and np.allclose(q[1, rows], q7(0.975), rtol=1e-9, atol=1e-12))
out = dict(stage="cfg5 full size, one GPU, device resident", d=d, m=m, nsim=N, n_test=n_test,
           model_and_setup_host_s=round(t_model, 2), sample_s=round(t_sample, 3), samples_per_s=round(N / t_sample),
           violations=int(st.violations), failed_chains=int(st.failed_chains), xi_gemm_s=round(t_xi, 3),
           paths_bands_s=round(t_bands, 2), gemm_ms=round(info["gemm_ms"], 1), bands_ms=round(info["bands_ms"], 1),
           gemm_tflops=round(info["gemm_tflops"], 2), slab_rows=info["slab_rows"],
           ysim_elements=float(n_test) * N, ysim_bytes_never_materialised=float(n_test) * N * 8,
           spot_check_3_rows_vs_sort=ok, bands_monotone_in_each_dim="n/a (random test points)",
           total_wall_s=round(time.perf_counter() - t_all, 1))
print(json.dumps(out), flush=True)
