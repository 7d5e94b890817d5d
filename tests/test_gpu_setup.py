"""Set-up stage on the device against numpy/LAPACK on the same inputs.

Tolerances (FP64, stated per test): a Cholesky factor is backward stable, so the checked
quantity is the reconstruction ||F F' - A|| / ||A|| <= 50 d eps; factors of well-conditioned
matrices are also compared entry-wise with LAPACK's at 1e-11 relative."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
EPS = np.finfo(float).eps


def _spd(rng, d, cond=1e3):
    Q, _ = np.linalg.qr(rng.standard_normal((d, d)))
    ev = np.logspace(0, -np.log10(cond), d)
    return (Q * ev) @ Q.T


@pytest.mark.parametrize("d", [1, 2, 63, 64, 65, 130, 500])
def test_chol_matches_lapack(lqg, d):
    from lineqgpr_b200 import setup_stage as S
    A = _spd(np.random.default_rng(d), d)
    Lf = S.chol(A)
    ref = np.linalg.cholesky((A + A.T) / 2)
    assert np.array_equal(np.triu(Lf, 1), np.zeros((d, d)))
    assert np.linalg.norm(Lf @ Lf.T - A) <= 50 * d * EPS * np.linalg.norm(A)
    assert np.max(np.abs(Lf - ref)) <= 1e-11 * np.max(np.abs(ref))


@pytest.mark.parametrize("d", [5, 64, 200, 333])
def test_whiten_box_device_is_tmgs_Ri(lqg, oracle, d):
    """U upper triangular, positive diagonal, U U' = Sigma; equals the oracle's literal tmg sequence
    (solve, chol, solve: tmg:R/tmg.R:15-28) to rounding on a well-conditioned Sigma."""
    from lineqgpr_b200 import setup_stage as S, samplers
    Sigma = _spd(np.random.default_rng(100 + d), d, cond=50.0)
    U = S.whiten_box_device(Sigma)
    assert np.array_equal(np.tril(U, -1), np.zeros((d, d)))
    assert np.all(np.diag(U) > 0)
    assert np.linalg.norm(U @ U.T - Sigma) <= 50 * d * EPS * np.linalg.norm(Sigma)
    assert np.max(np.abs(U - samplers.whiten_box(None, Sigma, "ul"))) <= 1e-12 * np.max(np.abs(U))
    Ri = oracle.whiten(np.linalg.inv(Sigma), np.zeros(d), np.zeros(d), np.zeros((0, d)), np.zeros(0))["Ri"] \
        if hasattr(oracle, "whiten") else samplers.whiten_box(None, Sigma, "tmg")
    assert np.max(np.abs(U - np.triu(Ri))) <= 1e-9 * np.max(np.abs(U))


def test_not_positive_definite_is_reported(lqg):
    from lineqgpr_b200 import setup_stage as S
    rng = np.random.default_rng(3)
    B = rng.standard_normal((150, 40))
    A = B @ B.T                                     # rank 40 < 150
    A[np.diag_indices(150)] -= 1e-3                 # indefinite
    assert not S.is_positive_definite(A)
    with pytest.raises(ValueError, match="positive definite"):
        S.whiten_box_device(A)
    with pytest.raises(ValueError):
        S.chol(A)
    assert S.is_positive_definite(A + 1.0 * np.eye(150))
    nanm = np.eye(70); nanm[69, 69] = np.nan
    assert not S.is_positive_definite(nanm)


@pytest.mark.parametrize("d,nrhs", [(3, 1), (64, 7), (129, 1), (300, 65)])
def test_solve_spd(lqg, d, nrhs):
    from lineqgpr_b200 import setup_stage as S
    rng = np.random.default_rng(d + nrhs)
    A = _spd(rng, d, cond=1e4)
    B = rng.standard_normal((d, nrhs))
    X = S.solve_spd(A, B)
    ref = np.linalg.solve(A, B)
    assert np.max(np.abs(X - ref)) <= 1e-10 * np.max(np.abs(ref))          # cond 1e4 x a few hundred eps


def test_eta_params_and_lambda_pinv_cfg1_cfg3(lqg):
    import importlib
    from lineqgpr_b200 import setup_stage as S
    sim_mod = importlib.import_module("lineqgpr_b200.simulate")   # (the package also exports a function of that name)
    rng = np.random.default_rng(5)
    for d, m in ((100, 100), (200, 100), (37, 20)):
        if d == m:
            Lam = np.tril(np.ones((m, m)))                                   # monotone-type square Lambda
        else:
            Lam = np.vstack([np.eye(m), np.diff(np.eye(m), axis=0), rng.standard_normal((d - 2 * m + 1, m))]) \
                if d >= 2 * m - 1 else rng.standard_normal((d, m))
        Sig = _spd(rng, m, 1e6)
        mu, xm = rng.standard_normal(m), rng.standard_normal(m)
        e_mu, e_map, Se = S.eta_params(Lam, mu, xm, Sig, 1e-5)
        assert np.allclose(e_mu, Lam @ mu, rtol=1e-13, atol=1e-13)
        assert np.allclose(e_map, Lam @ xm, rtol=1e-13, atol=1e-13)
        ref = Lam @ Sig @ Lam.T + 1e-5 * np.eye(d)
        assert np.max(np.abs(Se - ref)) <= 1e-13 * np.max(np.abs(ref)) * m
        Lp = S.lambda_pinv_device(Lam)
        ref = sim_mod.lambda_pinv(Lam)
        assert np.max(np.abs(Lp - ref)) <= 1e-11 * np.max(np.abs(ref)), (d, m)
        eta = Lam @ rng.standard_normal((m, 4))
        assert np.allclose(Lp @ eta, np.linalg.lstsq(Lam, eta, rcond=None)[0], rtol=1e-10, atol=1e-11)


def test_lambda_pinv_rank_deficient(lqg):
    from lineqgpr_b200 import setup_stage as S
    Lam = np.ones((10, 3))
    with pytest.raises(ValueError, match="rank"):
        S.lambda_pinv_device(Lam)


def test_hmc_box_whitens_on_device_when_U_is_null(lqg):
    """lqg_hmc_box(U = NULL) factors Sigma itself: same chains as with the host factor up to the
    rounding difference between the two factorizations (short run, well-conditioned Sigma)."""
    from lineqgpr_b200 import workloads as W
    par, _ = W.cfg3()
    par = dict(par, **{"class": "HMC"})
    ctl = {"burn.in": 2, "n.chains": 8, "seed": 11, "burn.in.exact": True}
    a = lqg.tmvrnorm(par, 64, dict(ctl))
    b = lqg.tmvrnorm(par, 64, dict(ctl, whiten="device"))
    assert lqg.last_stats["violations"] == 0
    assert np.all(b >= par["lb"][:, None]) and np.all(b <= par["ub"][:, None])
    # first emitted sample of every chain agrees closely (a handful of bounces after 3 transitions)
    first = np.arange(8) * 8
    assert np.max(np.abs(a[:, first] - b[:, first])) <= 1e-8


def test_cfg4_scale_setup(lqg):
    """d = 2000 (cfg4): the jittered, numerically near-singular Sigma_eta of the headline workload."""
    import time
    from lineqgpr_b200 import setup_stage as S, workloads as W, samplers
    par, aux = W.cfg4()
    Sigma = par["Sigma"]
    d = Sigma.shape[0]
    S.whiten_box_device(Sigma[:64, :64] + np.eye(64))          # warm-up (module load, pool)
    t0 = time.perf_counter(); U = S.whiten_box_device(Sigma); t_dev = time.perf_counter() - t0
    t0 = time.perf_counter(); Uh = samplers.whiten_box(None, Sigma, "ul"); t_host = time.perf_counter() - t0
    print(f"whiten d={d}: device {t_dev*1e3:.1f} ms (incl. 64 MB of PCIe copies), host LAPACK {t_host*1e3:.1f} ms")
    nS = np.linalg.norm(Sigma)
    assert np.linalg.norm(U @ U.T - Sigma) <= 50 * d * EPS * nS
    assert np.linalg.norm(U @ U.T - Sigma) <= 4 * max(np.linalg.norm(Uh @ Uh.T - Sigma), d * EPS * nS)
    Lam = aux["Lambda"]
    Lp = S.lambda_pinv_device(Lam)
    assert np.max(np.abs(Lp @ Lam - np.eye(Lam.shape[1]))) <= 1e-11


def _kkt_ok(D, dvec, A, b, x, tol):
    """x minimises 1/2 x'Dx - dvec'x s.t. Ax >= b iff the gradient D x - dvec is a non-negative
    combination of the active rows; checked by a non-negative least-squares fit on those rows."""
    import scipy.optimize as so
    r = A @ x - b
    assert r.min() >= -tol
    g = D @ x - dvec
    act = r <= 1e-6 * max(1.0, np.abs(b).max())
    if not act.any():
        return np.abs(g).max() <= 1e-6 * max(1.0, np.abs(dvec).max())
    lam, res = so.nnls(A[act].T, g)
    return res <= 1e-6 * max(1.0, np.linalg.norm(g))


@pytest.mark.parametrize("n,mc,seed", [(5, 8, 0), (40, 100, 1), (130, 70, 2), (64, 1, 3)])
def test_solve_qp_matches_host_ipm_and_kkt(lqg, n, mc, seed):
    from lineqgpr_b200 import setup_stage as S, model as M
    rng = np.random.default_rng(seed)
    D = _spd(rng, n, 1e3) * 10
    dvec = rng.standard_normal(n)
    A = rng.standard_normal((mc, n))
    b = A @ rng.standard_normal(n) - np.abs(rng.standard_normal(mc))      # strictly feasible point exists
    x, it = S.solve_qp_device(D, dvec, A, b, return_iters=True)
    ref = M.solve_qp(D, dvec, A, b)
    assert 0 < it < 60
    # both finish with the active-set polish: the exact minimiser, like quadprog
    assert np.max(np.abs(x - ref)) <= 1e-11 * max(1.0, np.abs(ref).max())
    assert (A @ x - b).min() >= -1e-11 * max(1.0, np.abs(b).max())
    assert _kkt_ok(D, dvec, A, b, x, 1e-10)
    # the interior-point iterates themselves (no polish) follow each other and are close to it
    xi = S.solve_qp_device(D, dvec, A, b, polish=False)
    refi = M.solve_qp(D, dvec, A, b, polish=False)
    assert np.max(np.abs(xi - refi)) <= 1e-7 * max(1.0, np.abs(refi).max())
    assert np.max(np.abs(xi - x)) <= 1e-4 * max(1.0, np.abs(x).max())


def test_solve_qp_unconstrained_and_errors(lqg):
    from lineqgpr_b200 import setup_stage as S, _lib
    rng = np.random.default_rng(9)
    D = _spd(rng, 30, 100.0)
    dvec = rng.standard_normal(30)
    x = S.solve_qp_device(D, dvec, np.zeros((0, 30)), np.zeros(0))
    assert np.allclose(x, np.linalg.solve(D, dvec), rtol=1e-10, atol=1e-12)
    with pytest.raises(_lib.LqgError, match="NOT_PD"):
        S.solve_qp_device(-D, dvec, np.zeros((0, 30)), np.zeros(0))
    with pytest.raises(_lib.LqgError, match="finite"):
        S.solve_qp_device(D, dvec, np.eye(30)[:2], np.array([0.0, -np.inf]))
    with pytest.raises(_lib.LqgError, match="BUDGET"):
        S.solve_qp_device(D, dvec, np.eye(30), np.full(30, 5.0), max_iter=2)


def test_map_qp_of_cfg1_and_cfg4(lqg):
    """The xi.map programme the reference solves in predict.lineqGP (R/lineqGP.R:524-531) at the
    cfg1 (m = 100, 99 rows) and cfg4 (m = 1000, 2999 rows) sizes, against the host IPM."""
    import time
    from lineqgpr_b200 import setup_stage as S, model as M, workloads as W
    for name, build in (("cfg1", W.cfg1), ("cfg4", W.cfg4)):
        par, aux = build()
        mdl = aux["model"] if "model" in aux else aux["lineq_model"]
        pred = M.predict_lineqGP(mdl, aux["xtest"], solve_map=False)
        Sigma = pred["Sigma"]
        inv = np.linalg.inv(Sigma); inv = (inv + inv.T) / 2
        dvec = pred["mu"] @ inv
        A, b = pred["model"]["lineqSys"]["M"], pred["model"]["lineqSys"]["g"]
        S.solve_qp_device(np.eye(4), np.ones(4), np.eye(4), np.zeros(4))            # warm-up
        t0 = time.perf_counter(); x, it = S.solve_qp_device(inv, dvec, A, b, return_iters=True); t_dev = time.perf_counter() - t0
        t0 = time.perf_counter(); ref = M.solve_qp(inv, dvec, A, b); t_host = time.perf_counter() - t0
        print(f"{name}: QP n={inv.shape[0]} rows={A.shape[0]}: device {t_dev*1e3:.0f} ms ({it} iterations), host IPM {t_host*1e3:.0f} ms")
        assert (A @ x - b).min() >= -1e-11
        x_ipm = S.solve_qp_device(inv, dvec, A, b, polish=False)
        assert not np.array_equal(x, x_ipm)                    # a consistent active set was found and solved
        assert np.max(np.abs(x - x_ipm)) <= 1e-5 * max(1.0, np.abs(x).max())
        f = lambda v: 0.5 * v @ inv @ v - dvec @ v
        assert f(x) <= f(ref) + 1e-9 * max(1.0, abs(f(ref)))
        # D = Sigma^-1 has cond ~ 1e10 at cfg4: the two exact-active-set solves agree to the accuracy
        # the Schur-complement systems allow
        assert np.max(np.abs(A @ (x - ref))) <= 1e-7 * max(1.0, np.abs(A @ ref).max())


def test_simulate_with_device_setup_matches_host_setup(lqg):
    """simulate.lineqGP / simulate.lineqAGP with control$setup = "device": the QP, the
    factorizations, eta-space parameters and Lambda^+ computed on the GPU give the same MAP and
    the same sample paths (to the accuracy two different factorizations allow) as the host set-up."""
    from lineqgpr_b200 import workloads as W
    for build in (W.cfg1, W.cfg3):
        par, aux = build()
        model = aux["model"]
        model = dict(model); model["localParam"] = dict(model["localParam"], sampler="HMC")
        ctl = {"n.chains": 16, "burn.in": 20}
        a = lqg.simulate(model, nsim=2000, seed=3, xtest=aux["xtest"], control=dict(ctl, setup="host"))
        b = lqg.simulate(model, nsim=2000, seed=3, xtest=aux["xtest"], control=dict(ctl, setup="device"))
        assert b["stats"]["violations"] == 0
        kmap = "xi.map" if "xi.map" in a else "xiAll.map"
        ksim = "xi.sim" if "xi.sim" in a else "xiAll.sim"
        assert np.max(np.abs(a[kmap] - b[kmap])) <= 1e-6 * max(1.0, np.abs(a[kmap]).max())
        sa, sb = a[ksim], b[ksim]
        assert sa.shape == sb.shape
        sd = sa.std(axis=1) + 1e-12
        # same seed, same Philox draws: the two runs follow each other until rounding differences in
        # the factor change a bounce decision; the path ensembles must agree well within MC error
        assert np.max(np.abs(sa.mean(axis=1) - sb.mean(axis=1)) / sd) <= 0.25
        assert np.max(np.abs(np.log(sb.std(axis=1) / sd))) <= 0.25


def test_basis_is_bit_identical_to_host_restatement(lqg):
    """lqg_basis vs basisCompute_lineqGP (R/lineqGP.R:29-60): equispaced and MaxMod-style irregular
    knots, points on knots, at both ends, ragged sizes; additive (block) form; out-of-range error."""
    from lineqgpr_b200 import setup_stage as S, model as M
    rng = np.random.default_rng(0)
    for m, n_t in ((2, 1), (10, 37), (100, 1000), (1000, 513)):
        u = np.linspace(0.0, 1.0, m)
        x = np.concatenate([rng.uniform(0, 1, n_t), u[: min(m, 5)], [0.0, 1.0]])
        assert np.array_equal(S.basis_device(x, u), M.basisCompute_lineqGP(x, u))
    u = np.array([0.0, 0.2047, 0.4020, 0.6275, 1.0])                      # cfg2's MaxMod knots
    x = rng.uniform(0, 1, 200)
    assert np.array_equal(S.basis_device(x, u), M.basisCompute_lineqGP(x, u))
    ulist = [np.linspace(0, 1, k) for k in (6, 4, 9)]
    X = rng.uniform(0, 1, (50, 3))
    ref = np.hstack([M.basisCompute_lineqGP(X[:, k], ulist[k]) for k in range(3)])   # R/lineqAGP.R:341-348
    assert np.array_equal(S.basis_device(X, ulist), ref)
    with pytest.raises(ValueError, match="knot range"):
        S.basis_device(np.array([0.5, 1.5]), np.linspace(0, 1, 5))


def test_paths_bands_with_device_basis(lqg):
    """paths_bands building each Phi slab on the device equals the run with a host-built Phi."""
    import importlib
    from lineqgpr_b200 import model as M
    sim_mod = importlib.import_module("lineqgpr_b200.simulate")
    rng = np.random.default_rng(1)
    ulist = [np.linspace(0, 1, 10) for _ in range(4)]
    X = rng.uniform(0, 1, (300, 4))
    xi = rng.standard_normal((40, 5000))
    Phi = np.hstack([M.basisCompute_lineqGP(X[:, k], ulist[k]) for k in range(4)])
    m0, q0, _ = sim_mod.paths_bands(Phi, xi, slab_rows=128)
    m1, q1, _ = sim_mod.paths_bands(None, xi, slab_rows=128, xtest=X, ulist=ulist)
    assert np.array_equal(m0, m1) and np.array_equal(q0, q1)


@pytest.mark.parametrize("d", [300, 1000])
def test_factorization_is_as_robust_as_lapack_on_nearly_singular_input(lqg, d):
    """Sigma_eta of this path is a rank-deficient product plus a tiny nugget (cfg4: cond ~ 1e10).
    The blocked factorisation must go through wherever LAPACK's does (shift >= 1e-11 ||S||) with
    the same backward error, and must refuse a matrix with a clearly negative direction."""
    from lineqgpr_b200 import setup_stage as S
    rng = np.random.default_rng(d)
    B = rng.standard_normal((d, d // 2))
    S0 = B @ B.T                                              # rank d/2
    nrm = np.linalg.norm(S0, 2)
    for shift in (1e-6, 1e-8, 1e-10, 1e-11):
        A = S0 + shift * nrm * np.eye(d)
        np.linalg.cholesky(A)                                 # LAPACK goes through
        U = S.whiten_box_device(A)
        Lf = S.chol(A)
        for F in (U, Lf):
            assert np.linalg.norm(F @ F.T - A) <= 50 * d * EPS * np.linalg.norm(A), shift
        X = S.solve_spd(A, A @ np.ones((d, 2)))
        assert np.linalg.norm(A @ X - A @ np.ones((d, 2))) <= 1e-9 * np.linalg.norm(A @ np.ones((d, 2))), shift
    assert not S.is_positive_definite(S0 - 1e-8 * nrm * np.eye(d))


def test_tmvrnorm_default_whitens_on_device_and_reports_non_pd(lqg):
    from lineqgpr_b200 import workloads as W
    par, _ = W.cfg3()
    par = dict(par, **{"class": "HMC"})
    a = lqg.tmvrnorm(par, 32, {"burn.in": 2, "n.chains": 4, "seed": 5})                      # default: device
    b = lqg.tmvrnorm(par, 32, {"burn.in": 2, "n.chains": 4, "seed": 5, "whiten": "device"})
    assert np.array_equal(a, b)
    S_bad = np.eye(par["mu"].size)
    S_bad[0, 1] = S_bad[1, 0] = 2.0                                                          # positive diagonal, eigenvalues 1 +- 2
    bad = dict(par, Sigma=S_bad)
    with pytest.raises(ValueError, match="positive definite"):
        lqg.tmvrnorm(bad, 8, {"burn.in": 1, "n.chains": 2, "seed": 1})


def test_pageable_and_pinned_results_are_bit_identical(lqg):
    """A pageable result matrix goes through the pinned staging ring and helper threads, a
    page-locked one straight through the copy engine: same bytes, ragged chain counts and a
    sample count that does not fill the last round."""
    from lineqgpr_b200 import workloads as W, _lib as L
    par, _ = W.cfg4(m=300)
    par = dict(par, **{"class": "HMC"})
    d = par["mu"].size
    for n_chains, nsim in ((7, 7 * 150), (32, 32 * 41)):
        ctl = {"burn.in": 3, "n.chains": n_chains, "seed": 2}
        x_page = lqg.tmvrnorm(par, nsim, dict(ctl))
        out = L.empty_result(d, nsim)
        x_pin = lqg.tmvrnorm(par, nsim, dict(ctl, out=out))
        assert x_page.shape == x_pin.shape == (d, nsim)
        assert np.array_equal(x_page, x_pin)
        assert np.all(x_page >= par["lb"][:, None]) and np.all(x_page <= par["ub"][:, None])


def test_large_dgemm_result_through_staging_ring(lqg):
    """A > 32 MB pageable GEMM result is delivered through the pinned staging ring (16 MB pieces,
    helper threads): every byte must arrive, including the ragged last piece."""
    rng = np.random.default_rng(4)
    A = rng.standard_normal((701, 64))
    B = rng.standard_normal((64, 9001))                     # C = 701 x 9001 doubles = 50.5 MB, not a multiple of 16 MB
    C = lqg.dgemm(A, B)
    assert np.allclose(C, A @ B, rtol=1e-12, atol=1e-12)
    assert np.array_equal(C, lqg.dgemm(A, B))
