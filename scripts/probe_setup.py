"""Times the device set-up stage at cfg4 scale (d = 2000, m = 1000) next to numpy/LAPACK."""
import ctypes as C
import sys
import time
import numpy as np
sys.path.insert(0, ".")
import importlib
from lineqgpr_b200 import _lib as L, workloads as W, setup_stage as S, samplers
sim_mod = importlib.import_module("lineqgpr_b200.simulate")

par, aux = W.cfg4()
Sigma, Lam = L.f64(par["Sigma"]), L.f64(aux["Lambda"])
d, m = Lam.shape
U = np.empty((d, d), order="F")
ms = C.c_double(0)
for rep in range(3):
    t0 = time.perf_counter()
    L.check(L.lib().lqg_whiten_box(d, L.ptr(Sigma), L.make_opts(), L.ptr(U), C.byref(ms)), "whiten")
    wall = time.perf_counter() - t0
    print(f"lqg_whiten_box d={d}: kernels {ms.value:.2f} ms, call wall (pageable host in/out) {wall*1e3:.1f} ms")
t0 = time.perf_counter(); samplers.whiten_box(None, Sigma, "ul"); print(f"numpy UL factor: {(time.perf_counter()-t0)*1e3:.1f} ms")
pm = aux["model"] if "model" in aux else None
rng = np.random.default_rng(0)
Sx = Sigma[:m, :m].copy()
for rep in range(2):
    t0 = time.perf_counter(); S.eta_params(Lam, np.zeros(m), np.zeros(m), Sx, 1e-7); print(f"eta_params device call: {(time.perf_counter()-t0)*1e3:.1f} ms")
t0 = time.perf_counter(); ref = Lam @ Sx @ Lam.T; print(f"numpy Lambda Sigma Lambda': {(time.perf_counter()-t0)*1e3:.1f} ms")
for rep in range(2):
    t0 = time.perf_counter(); Lp = S.lambda_pinv_device(Lam); print(f"lambda_pinv device call: {(time.perf_counter()-t0)*1e3:.1f} ms")
t0 = time.perf_counter(); sim_mod.lambda_pinv(Lam); print(f"numpy QR pinv: {(time.perf_counter()-t0)*1e3:.1f} ms")
for rep in range(2):
    t0 = time.perf_counter(); ok = S.is_positive_definite(Sigma); print(f"is_positive_definite device call: {(time.perf_counter()-t0)*1e3:.1f} ms -> {ok}")
t0 = time.perf_counter(); np.linalg.eigvalsh(Sigma); print(f"numpy eigvalsh: {(time.perf_counter()-t0)*1e3:.1f} ms")
