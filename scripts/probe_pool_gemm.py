"""The momentum-pool GEMM alone at the cfg4 shape: X = U A, U upper triangular 2000 x 2000, A 2000 x (296*64).
usage: python scripts/probe_pool_gemm.py   (LQG_GEMM_TAIL=<waves> selects how many trailing column blocks go longest-first)"""
import ctypes as C, json, os, sys
import numpy as np
sys.path.insert(0, ".")
import torch
from lineqgpr_b200 import _lib as L
lib = L.lib()
dev = torch.device("cuda", 0)
d, n = 2000, 296 * 64
g = torch.Generator(device=dev); g.manual_seed(1)
U = torch.triu(torch.randn((d, d), dtype=torch.float64, device=dev, generator=g)).T.contiguous()   # column-major upper triangular
A = torch.randn((n, d), dtype=torch.float64, device=dev, generator=g)
Cm = torch.empty((n, d), dtype=torch.float64, device=dev)
o = L.make_opts(mem=L.MEM_DEVICE, stream=torch.cuda.current_stream().cuda_stream)
ms = C.c_double(0)
dfma, dmma = C.c_double(), C.c_double()
L.check(lib.lqg_measure_fp64_peaks(C.byref(dfma), C.byref(dmma), None), "peaks")
# the public lqg_dgemm has no triangular flag: time the general product (full 2 d^2 n flop) here; the pool's
# triangular launch is timed by bench.py (roofline_prepass)
best = 1e9
for rep in range(5):
    L.check(lib.lqg_dgemm(d, n, d, U.data_ptr(), d, A.data_ptr(), d, Cm.data_ptr(), d, None, o, C.byref(ms)), "dgemm")
    best = min(best, ms.value)
print(json.dumps({"shape": [d, n, d], "ms": best, "tflops": 2.0 * d * d * n / (best * 1e-3) / 1e12, "dmma_peak": dmma.value}))
