"""One cfg5-sized slab of the path stage through lqg_paths (host arrays in, bands out): 4224 test points x 2000
basis functions x 125 208 samples, dense Phi -> pilot product, fused product (dgemm_tma_kernel<true>), reduce, select.
With LQG_BANDS_FUSED=0 the same call stores the slab and runs the band kernels on it.
ncu:  ncu --set full --clock-control none --import-source on -k regex:dgemm_tma -c 2 -o gpurun_out/fused python scripts/profile_fused_bands.py
usage: python scripts/profile_fused_bands.py [n_t] [m] [N]"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import lineqgpr_b200 as lqg

n_t = int(sys.argv[1]) if len(sys.argv) > 1 else 4224
m = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
N = int(sys.argv[3]) if len(sys.argv) > 3 else 125208
rng = np.random.default_rng(0)
Phi = rng.uniform(size=(n_t, m))
xi = rng.standard_normal((m, N), dtype=np.float32).astype(np.float64)
probs = np.array([0.025, 0.975])
for rep in range(2):
    t = time.perf_counter()
    mean, q, info = lqg.paths(Phi, xi, probs)
    wall = time.perf_counter() - t
    print(json.dumps(dict(stage="lqg_paths", n_t=n_t, m=m, N=N, wall_s=round(wall, 3),
                          **{k: (round(v, 3) if isinstance(v, float) else v) for k, v in info.items()},
                          product_tflops=round(2.0 * n_t * m * N / (info["gemm_ms"] * 1e-3) / 1e12, 2))), flush=True)
