"""Tiny end-to-end pass over every kernel family, for compute-sanitizer (memcheck / racecheck)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import lineqgpr_b200 as lqg
from lineqgpr_b200 import workloads as W

rng = np.random.default_rng(0)
# box kernel variants: d = 10 (1 warp), 100 (1 warp x 4), 200 (2 warps), 300 (4 warps), 1100 (8 warps, smem thresholds)
for name, par, n, ch in (("cfg3", W.cfg3()[0], 12, 4), ("cfg1", W.cfg1()[0], 8, 4), ("cfg4_m100", W.cfg4(m=100)[0], 6, 3),
                         ("cfg5_D30", W.cfg5(D=30, knots=10, n=150, n_test=10)[0], 6, 3), ("cfg4_m550", W.cfg4(m=550)[0], 4, 2)):
    for prepass in (0, 1):
        x = lqg.tmvrnorm(dict(par, **{"class": "HMC"}), n, {"burn.in": 1, "n.chains": ch, "seed": 1, "prepass": prepass})
        assert lqg.last_stats["violations"] == 0, name
    print("box", name, x.shape, flush=True)
# canonical kernel: F in shared memory (TMA bulk copy) and streamed
F = rng.standard_normal((8, 10)); z0 = 0.1 * rng.standard_normal(10); g = 0.5 + np.maximum(0, -(F @ z0))
z, st = lqg.rtmg_canonical(5, 3, z0, F, g, n_chains=3, burn_in=1); print("canonical small", z.shape, st["bounces"], flush=True)
F = rng.standard_normal((300, 120)); z0 = np.zeros(120); g = np.full(300, 6.0)
z, st = lqg.rtmg_canonical(2, 3, z0, F, g, n_chains=2, max_restarts=20000); print("canonical streamed", z.shape, st["bounces"], flush=True)
# RSM, ExpT
par, _ = W.cfg2(sampler="RSM")
x = lqg.tmvrnorm(par, 50, {"seed": 2}); print("rsm", x.shape, flush=True)
x = lqg.tmvrnorm(dict(W.cfg3()[0], **{"class": "ExpT"}), 300, {"seed": 2}); print("expt", x.shape, flush=True)
# GEMM with ragged sizes + row sums, bands with ties
A = rng.standard_normal((130, 19)); B = rng.standard_normal((19, 67))
C, rs = lqg.dgemm(A, B, row_sum=True); assert np.allclose(C, A @ B)
y = rng.standard_normal((5, 300)); y[0, :100] = y[0, 0]
m, q = lqg.bands(y, (0.025, 0.5, 0.975))
yb = rng.standard_normal((37, 9001)); yb[3, :5000] = 1.0          # sampled (bracketed) bands path + one overflowing row
mb, qb = lqg.bands(yb, (0.0, 0.3, 1.0)); assert np.allclose(qb[1], np.quantile(yb, 0.3, axis=1), rtol=1e-12, atol=1e-12)
print("gemm/bands ok", flush=True)
# set-up stage: blocked Cholesky / solves / QP at ragged sizes (1, 64 +- 1, several blocks)
from lineqgpr_b200 import setup_stage as S
for n in (1, 63, 65, 130):
    Q = rng.standard_normal((n, n)); A = Q @ Q.T + n * np.eye(n)
    Lf = S.chol(A); assert np.allclose(Lf @ Lf.T, A)
    Uf = S.whiten_box_device(A); assert np.allclose(Uf @ Uf.T, A)
    X = S.solve_spd(A, rng.standard_normal((n, 3)))
    Lam = rng.standard_normal((n + 7, n)); Lp = S.lambda_pinv_device(Lam); assert np.allclose(Lp @ Lam, np.eye(n), atol=1e-8)
    S.eta_params(Lam, np.ones(n), np.ones(n), A, 1e-6)
    Aq = rng.standard_normal((2 * n + 1, n)); bq = Aq @ rng.standard_normal(n) - 1.0
    S.solve_qp_device(A, rng.standard_normal(n), Aq, bq)
x = lqg.tmvrnorm(dict(W.cfg3()[0], **{"class": "HMC"}), 16, {"burn.in": 1, "n.chains": 4, "seed": 1, "whiten": "device"})
print("setup stage ok", flush=True)
from lineqgpr_b200 import _lib
import os
print("guard mode", os.environ.get("LQG_GUARD", "0"), "violations", _lib.lib().lqg_guard_violations(), flush=True)
assert _lib.lib().lqg_guard_violations() == 0
