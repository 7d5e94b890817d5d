"""Generates tests/golden/*.npz.  Run in the build container (needs /root/reference):

    make -C oracle && make -C oracle ref && python tests/golden/make_golden.py

Every `z_ref` array below is the output of the UNMODIFIED reference chain
(tmg/src/HmcSampler.cpp compiled into oracle/_ref/libtmg_ref.so, see oracle/Makefile) for the
stored inputs and seed; `draws` is the Gaussian stream that build consumes for that seed
(libstdc++ knuth_b + normal_distribution, tmg/src/HmcSampler.h:50-52), so a GPU run fed the
same draws must reproduce `z_ref`.  RSM / band fixtures come from the numpy restatement of the
R glue (no R in the image: parity unpinned for those, see oracle/lineqgpr_oracle.py)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import lineqgpr_oracle as O  # noqa: E402
from lineqgpr_b200 import workloads as W  # noqa: E402

OUT = Path(__file__).resolve().parent


def canonical_case(name, F, g, z0, n, seed):
    z_ref = O.ref_rtmg_canonical(n, seed, z0, F, g)
    z_or, info = O.rtmg_canonical(n, seed, z0, F, g)
    assert np.array_equal(z_ref, z_or), name
    draws = O.normal_stream(seed, info["draws"])
    np.savez_compressed(OUT / f"{name}.npz", F=F, g=g, z0=z0, n=n, seed=seed, z_ref=z_ref, draws=draws,
                        bounces=info["bounces"], vel_restarts=info["vel_restarts"], chk_restarts=info["chk_restarts"])
    print(name, z_ref.shape, {k: info[k] for k in ("bounces", "vel_restarts", "chk_restarts", "draws")})


def main():
    O.build(ref=True)
    # (1) tmg's own Rd example restricted to its two linear constraints (tmg/man/rtmg.Rd:82-91):
    #     x >= 0 on both coordinates after whitening of M = [[.5,-.4],[-.4,.5]], r = 0, initial (4,1)
    M = np.array([[0.5, -0.4], [-0.4, 0.5]])
    w = O.whiten(M, np.zeros(2), np.array([4.0, 1.0]), np.eye(2), np.zeros(2))
    canonical_case("rd_example_2d", w["f2"], w["g2"], w["initial2"], 200, 4711)
    np.savez_compressed(OUT / "rd_example_2d_frame.npz", Ri=w["Ri"], Mir=w["Mir"], M=M)

    # (2) random dense polytope, d = 10, c = 8
    rng = np.random.default_rng(11)
    F = rng.standard_normal((8, 10)); z0 = 0.1 * rng.standard_normal(10)
    g = 0.5 * np.abs(rng.standard_normal(8)) + np.maximum(0, -(F @ z0)) + 0.05
    canonical_case("dense_d10_c8", F, g, z0, 100, 1234)

    # (3) roxygen example shape (R/lineqGPsamplers.R:86-91): 100-dim Gaussian-kernel Sigma + 1e-9 I,
    #     box [-1,1] -> canonical frame through the tmvrnorm.HMC glue; 5 burn-in + 10 samples
    x = np.linspace(0, 1, 100)
    Sigma = np.exp(-0.5 * np.subtract.outer(x / 0.2, x / 0.2) ** 2) + 1e-9 * np.eye(100)
    H, r, init, f, gg = O.hmc_setup(np.zeros(100), Sigma, -np.ones(100), np.ones(100))
    w = O.whiten(H, r, init, f, gg)
    canonical_case("roxygen_box_d100", w["f2"], w["g2"], w["initial2"], 15, 99)

    # (4) cfg1 (1-D monotone, m = 100) in the canonical frame, 30 transitions
    par, _ = W.cfg1()
    H, r, init, f, gg = O.hmc_setup(par["mu"], par["Sigma"], par["lb"], par["ub"], par["map"])
    w = O.whiten(H, r, init, f, gg)
    canonical_case("cfg1_canonical", w["f2"], w["g2"], w["initial2"], 30, 2024)
    np.savez_compressed(OUT / "cfg1_frame.npz", Ri=w["Ri"], Mir=w["Mir"], init=init, mu=par["mu"],
                        Sigma=par["Sigma"], lb=par["lb"], ub=par["ub"])

    # (5) a stiff case with velocity and feasibility restarts (d = 60, c = 200)
    rng = np.random.default_rng(5)
    F = rng.standard_normal((200, 60)); z0 = 0.1 * rng.standard_normal(60)
    g = 0.5 * np.abs(rng.standard_normal(200)) + np.maximum(0, -(F @ z0)) + 0.05
    canonical_case("stiff_d60_c200", F, g, z0, 12, 1234)

    # (6) RSM on cfg2 with injected draws
    par, _ = W.cfg2(sampler="RSM")
    rng = np.random.default_rng(21)
    Z = rng.standard_normal((20, 6000)); U = rng.uniform(size=3000)
    xi, info = O.tmvrnorm_RSM(par, 200, Z, U)
    np.savez_compressed(OUT / "rsm_cfg2.npz", Z=Z, U=U, xi=xi, accepted_idx=info["accepted_idx"],
                        inbox=info["inbox"], n_proposals=info["n_proposals"], factor=info["factor"], w=info["w"],
                        map=par["map"], lb=par["lb"], ub=par["ub"])
    print("rsm_cfg2", xi.shape, info["n_proposals"], int(info["inbox"].sum()))

    # (7) band summaries
    rng = np.random.default_rng(3)
    y = rng.standard_normal((7, 1001)) * np.arange(1, 8)[:, None]
    mean, q = O.bands(y, (0.025, 0.5, 0.975))
    np.savez_compressed(OUT / "bands.npz", y=y, mean=mean, q=q, probs=np.array([0.025, 0.5, 0.975]))


if __name__ == "__main__":
    main()
