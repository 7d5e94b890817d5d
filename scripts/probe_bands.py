"""Bands kernel alone on a device-resident slab: bracketed path vs exact path (LQG_BANDS_EXACT=1)."""
import ctypes as C
import os, sys, json
import numpy as np
sys.path.insert(0, ".")
import torch
from lineqgpr_b200 import _lib as L
lib = L.lib()
dev = torch.device("cuda", 0)
for (n_t, N) in ((2000, 200000), (4000, 100000), (500, 1000000)):
    g = torch.Generator(device=dev); g.manual_seed(1)
    y = torch.randn((N, n_t), dtype=torch.float64, device=dev, generator=g)          # column-major n_t x N
    y = y * torch.linspace(0.1, 10.0, n_t, dtype=torch.float64, device=dev)[None, :] + 0.3
    probs = torch.tensor([0.025, 0.5, 0.975], dtype=torch.float64, device=dev)
    mean = torch.empty(n_t, dtype=torch.float64, device=dev); q = torch.empty((n_t, 3), dtype=torch.float64, device=dev)
    o = L.make_opts(mem=L.MEM_DEVICE, stream=torch.cuda.current_stream().cuda_stream)
    ms = C.c_double(0)
    best = 1e9
    for rep in range(3):
        L.check(lib.lqg_bands(n_t, N, y.data_ptr(), n_t, probs.data_ptr(), 3, mean.data_ptr(), q.data_ptr(), o, C.byref(ms)), "bands")
        best = min(best, ms.value)
    gb = n_t * N * 8 / 1e9
    ref = torch.quantile(y[:, :4].T.contiguous(), probs, dim=1).T     # spot check of 4 rows (torch: linear interpolation = type 7)
    ok = bool(torch.allclose(ref, q[:4], rtol=1e-12, atol=1e-12))
    print(json.dumps(dict(stage="bands", path="exact" if os.environ.get("LQG_BANDS_EXACT") == "1" else "bracketed", n_t=n_t, N=N, slab_GB=round(gb, 2),
                          ms=round(best, 3), GB_per_s_algorithmic=round(gb / (best * 1e-3)), spot_check=ok)), flush=True)
    del y
