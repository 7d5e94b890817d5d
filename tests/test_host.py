"""CPU tests: host logic, C-ABI surface, loud failure without a GPU."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def test_header_and_library_export_the_same_symbols(lqg):
    from lineqgpr_b200 import _lib
    header = (ROOT / "include" / "lineqgpr_b200.h").read_text()
    declared = set(re.findall(r"\b(lqg_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS)
    L = ctypes.CDLL(str(_lib.LIB_PATH))
    for s in declared:
        assert hasattr(L, s), s
    assert b"sm_100a" in _lib.lib().lqg_version()


def test_struct_layouts_match_header(lqg):
    from lineqgpr_b200 import _lib
    assert ctypes.sizeof(_lib.Opts) == 88
    assert ctypes.sizeof(_lib.HmcStats) == 144
    assert ctypes.sizeof(_lib.RsmStats) == 40
    assert ctypes.sizeof(_lib.ExptStats) == 32


def test_no_gpu_fails_loudly(lqg):
    """No CPU fallback: without a device the product path raises, it never computes on the host."""
    from lineqgpr_b200 import _lib
    if _lib.lib().lqg_device_count() > 0:
        pytest.skip("a GPU is present")
    obj = dict(mu=np.zeros(3), Sigma=np.eye(3), lb=-np.ones(3), ub=np.ones(3))
    with pytest.raises(_lib.LqgError) as e:
        lqg.tmvrnorm(obj, 5, {"burn.in": 1}, method="HMC")
    assert e.value.status == 4 and "no CPU fallback" in str(e.value)
    with pytest.raises(_lib.LqgError):
        lqg.tmvrnorm(obj, 5, method="RSM")
    with pytest.raises(_lib.LqgError):
        lqg.dgemm(np.eye(2), np.eye(2))
    # the set-up stage has no host fallback either: "device" means device
    from lineqgpr_b200 import setup_stage as S
    for call in (lambda: S.whiten_box_device(np.eye(3)), lambda: S.chol(np.eye(3)),
                 lambda: S.is_positive_definite(np.eye(3)), lambda: S.solve_spd(np.eye(3), np.ones(3)),
                 lambda: S.eta_params(np.eye(3), np.ones(3), np.ones(3), np.eye(3)),
                 lambda: S.lambda_pinv_device(np.eye(3)), lambda: S.basis_device(np.array([0.5]), np.linspace(0, 1, 3)),
                 lambda: S.solve_qp_device(np.eye(3), np.ones(3), np.eye(3), np.zeros(3))):
        with pytest.raises(_lib.LqgError) as e:
            call()
        assert e.value.status == 4
    with pytest.raises(_lib.LqgError):
        lqg.tmvrnorm(obj, 5, {"burn.in": 1, "whiten": "device"}, method="HMC")


def test_setup_mode_is_explicit(lqg):
    """predict/simulate take setup = "host" | "device"; anything else is an error, and the host
    algebra provider is plain numpy (usable without a GPU)."""
    from lineqgpr_b200 import model as M
    with pytest.raises(ValueError, match="host \\| device"):
        M._algebra("auto")
    S = np.array([[2.0, 0.5], [0.5, 1.0]])
    assert np.allclose(M._algebra("host").chol2inv(S) @ S, np.eye(2))
    assert not M._algebra("host").min_eig_nonpositive(S)
    assert M._algebra("host").min_eig_nonpositive(S - 3 * np.eye(2))


def test_product_package_never_imports_the_oracle():
    for p in (ROOT / "lineqgpr_b200").rglob("*.py"):
        assert "oracle" not in p.read_text().lower(), p
    for p in (ROOT / "lineqgpr_b200" / "csrc").glob("*.cu*"):
        assert "oracle" not in p.read_text().lower(), p


def test_tmvrnorm_dispatch_and_argument_errors(lqg):
    obj = dict(mu=np.zeros(2), Sigma=np.eye(2), lb=np.zeros(2), ub=np.ones(2))
    with pytest.raises(TypeError):
        lqg.tmvrnorm(obj, 3)                                   # no class
    with pytest.raises(TypeError):
        lqg.tmvrnorm(dict(obj, **{"class": "Gibbs"}), 3)          # removed after lineqGPR 0.0.3
    bad = dict(obj, lb=np.array([0.0, 2.0]))
    with pytest.raises(ValueError, match='has to be greater'):
        lqg.tmvrnorm(bad, 3, method="HMC")
    with pytest.raises(ValueError, match="d-dimensionals"):
        lqg.tmvrnorm(dict(obj, lb=np.zeros(3)), 3, method="HMC")


def test_hat_basis_and_systems(lqg):
    u = np.linspace(0, 1, 5)
    x = np.array([0.0, 0.1, 0.25, 0.99, 1.0])
    Phi = lqg.basisCompute_lineqGP(x, u)
    assert np.allclose(Phi.sum(1), 1.0) and np.allclose(Phi @ u, x)      # partition of unity, linear precision
    assert (Phi > 0).sum(1).max() <= 2
    with pytest.raises(ValueError):
        lqg.basisCompute_lineqGP(np.array([1.2]), u)
    s = lqg.lineqGPSys(5, "monotonicity", 0, np.inf, lineqSysType="oneside")
    assert s["M"].shape == (4, 5) and np.allclose(s["M"] @ np.arange(5.0), 1.0)
    s2 = lqg.lineqGPSys(5, "monotonicity", 0, np.inf, rmInf=False)
    assert s2["A"].shape == (5, 5) and s2["l"][0] == -np.inf and np.all(s2["l"][1:] == 0)
    s3 = lqg.lineqGPSys([3, 2], "monotonicity", 0, np.inf, d=2, rmInf=False)
    assert s3["A"].shape == (12, 6)
    s4 = lqg.lineqGPSys(4, "convexity", 0, np.inf, lineqSysType="oneside")
    assert s4["M"].shape == (2, 4) and np.allclose(s4["M"] @ (np.arange(4.0) ** 2), 2.0)


def test_config_shapes_match_survey_table(lqg):
    from lineqgpr_b200 import workloads as W
    expect = {"cfg1": (100, 100, 99), "cfg2": (10, 20, 13), "cfg3": (10, 10, 8)}
    for name, (m, d, c) in expect.items():
        par, aux = W.CONFIGS[name]()
        assert aux["Lambda"].shape == (d, m)
        assert par["Sigma"].shape == (d, d)
        assert int(np.isfinite(par["lb"]).sum() + np.isfinite(par["ub"]).sum()) == c
        init = np.minimum(np.maximum(par.get("map", par["mu"]), par["lb"] + 1e-9), par["ub"] - 1e-9)
        assert np.all(init > par["lb"]) and np.all(init < par["ub"])
    par, aux = W.cfg4(m=60)
    assert aux["Lambda"].shape == (120, 60) and int(np.isfinite(par["lb"]).sum() + np.isfinite(par["ub"]).sum()) == 179
    par, aux = W.cfg5(D=6, knots=5, n=60, n_test=50)
    assert aux["Lambda"].shape == (30, 30) and int(np.isfinite(par["lb"]).sum()) == 24


def test_qp_solver_kkt(lqg):
    from lineqgpr_b200.model import solve_qp
    rng = np.random.default_rng(0)
    B = rng.standard_normal((15, 15)); D = B @ B.T + 0.1 * np.eye(15); dv = rng.standard_normal(15)
    A = np.vstack([np.eye(15), -np.eye(15)]); b = np.concatenate([np.full(15, -0.1), np.full(15, -0.3)])
    x = solve_qp(D, dv, A, b)
    s = A @ x - b
    assert s.min() > -1e-9
    grad = D @ x - dv                      # = A' lam with lam >= 0 supported on the active set
    active = s < 1e-6
    lam, *_ = np.linalg.lstsq(A[active].T, grad, rcond=None)
    assert np.allclose(A[active].T @ lam, grad, atol=1e-6) and lam.min() > -1e-6


def test_lambda_pinv_is_qr_solve(lqg, oracle):
    from lineqgpr_b200.simulate import lambda_pinv
    rng = np.random.default_rng(4)
    sq = lqg.lineqGPSys(6, "monotonicity", 0, np.inf, rmInf=False)["A"]
    tall = np.vstack([np.eye(6), sq])
    for Lam in (sq, tall):
        eta = rng.standard_normal((Lam.shape[0], 5))
        assert np.allclose(lambda_pinv(Lam) @ eta, oracle.qr_solve(Lam, eta), atol=1e-12)


def test_whitening_factor_ul_equals_tmg_sequence(lqg):
    """The one-Cholesky 'ul' factor is the matrix tmg's inverse/chol/inverse sequence produces."""
    from lineqgpr_b200.samplers import whiten_box
    from lineqgpr_b200 import workloads as W
    rng = np.random.default_rng(6)
    B = rng.standard_normal((30, 30)); S = B @ B.T + 0.5 * np.eye(30)
    for Sigma in (S, W.cfg1()[0]["Sigma"]):
        Uu = whiten_box(None, Sigma, "ul"); Ut = whiten_box(None, Sigma, "tmg")
        assert np.allclose(np.tril(Uu, -1), 0) and np.all(np.diag(Uu) > 0)
        assert np.allclose(Uu @ Uu.T, Sigma, rtol=1e-10, atol=1e-12 * np.abs(Sigma).max())
        # tmg's sequence inverts Sigma twice: agreement is limited by cond(Sigma) * eps
        scale = np.abs(Uu).max()
        assert np.abs(Uu - Ut).max() / scale < 1e-16 * np.linalg.cond(Sigma) * 100
    with pytest.raises(ValueError, match="positive definite"):
        whiten_box(None, np.array([[1.0, 2.0], [2.0, 1.0]]), "ul")
    with pytest.raises(ValueError, match="positive definite"):
        whiten_box(None, np.array([[1.0, 2.0], [2.0, 1.0]]), "tmg")


def test_expt_setup_anchors(lqg):
    """Independent case: minimax tilting is exact, psi* = log P(l < X < u); the two independently
    written set-ups (product / oracle) agree."""
    from scipy.stats import norm
    from lineqgpr_b200 import expt
    from oracle import expt_oracle as EO
    l = np.array([0.5, -1.0, -np.inf]); u = np.array([2.0, 0.3, 1.0])
    su = expt.setup(l, u, np.eye(3))
    exact = np.log((norm.cdf(2) - norm.cdf(.5)) * (norm.cdf(.3) - norm.cdf(-1)) * norm.cdf(1.0))
    assert abs(su["psistar"] - exact) < 1e-12
    rng = np.random.default_rng(0)
    B = rng.standard_normal((5, 5)); S = B @ B.T + 0.3 * np.eye(5)
    l = -rng.uniform(0.1, 1, 5); u = rng.uniform(0.1, 2, 5)
    a, b = expt.setup(l, u, S), EO.setup(l, u, S)
    assert np.array_equal(a["perm"], b["perm"]) and abs(a["psistar"] - b["psistar"]) < 1e-9
    assert np.abs(a["L0"] - b["L0"]).max() < 1e-12 and np.abs(a["mu"] - b["mu"]).max() < 1e-7
    with pytest.raises(ValueError):
        expt.setup(np.array([1.0]), np.array([0.0]), np.eye(1))


def test_solve_qp_stops_at_the_rounding_floor(lqg):
    """A problem where the merit bottoms out just above tol (the dual residual floor grows like
    1/mu in the normal-equations form): the best iterate is returned instead of iterating to
    max_iter while the residual drifts."""
    from lineqgpr_b200 import model as M
    rng = np.random.default_rng(1)
    for n in (1, 63, 65, 130):                              # the n = 130 draw of this stream used to stall
        Q = rng.standard_normal((n, n)); D = Q @ Q.T + n * np.eye(n)
        A = rng.standard_normal((2 * n + 1, n)); b = A @ rng.standard_normal(n) - 1.0
        dvec = rng.standard_normal(n)
        x = M.solve_qp(D, dvec, A, b)
        x_ref = M.solve_qp(D, dvec, A, b, tol=1e-9)
        assert (A @ x - b).min() >= -1e-9
        assert np.max(np.abs(x - x_ref)) <= 1e-7 * max(1.0, np.abs(x_ref).max())
        g = D @ x - dvec                                     # stationarity on the active set
        act = (A @ x - b) < 1e-7
        lam = np.linalg.lstsq(A[act].T, g, rcond=None)[0] if act.any() else np.zeros(0)
        assert np.linalg.norm(A[act].T @ lam - g) <= 1e-6 * max(1.0, np.linalg.norm(g))
        assert lam.min(initial=0.0) >= -1e-6


def test_r_shim_type_checks_against_the_header():
    """R is not installed here, so rshim/lqg_rcall.c cannot be built; it is at least compiled
    (syntax + types, -Wall -Wextra clean) against include/lineqgpr_b200.h with a minimal stand-in
    for R's C API, so that every .Call wrapper matches the C ABI it forwards to."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    r = subprocess.run([gcc, "-fsyntax-only", "-Wall", "-Wextra", "-Werror", f"-I{ROOT / 'include'}",
                        f"-I{ROOT / 'tests' / 'r_stub'}", str(ROOT / "rshim" / "lqg_rcall.c")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # every wrapper the patched R methods call exists in the shim
    shim = (ROOT / "rshim" / "lqg_rcall.c").read_text()
    rfile = (ROOT / "rshim" / "lineqGPsamplers_b200.R").read_text()
    for name in set(re.findall(r'\.Call\("(lqg_[a-z_]+)"', rfile)):
        assert re.search(r"\bSEXP %s\(" % name, shim), name


def test_fork_defaults_do_not_depend_on_the_shard():
    """Families of chains sharing a root chain's burn-in (lqg_opts.fork): the wrapper's default must be the same
    for a whole job and for any shard of it (a shard is recognisable by its chain.offset), a lone chain without
    an offset is the reference's single serial chain, and injected draws never fork."""
    from lineqgpr_b200.samplers import chain_fork, chain_burn_in, FORK_DEFAULT, FORK_BURN_IN
    assert chain_fork(444) == (FORK_DEFAULT, FORK_BURN_IN)
    assert chain_fork(2) == (FORK_DEFAULT, FORK_BURN_IN)
    assert chain_fork(1) == (0, 0)
    assert chain_fork(1, {"chain.offset": 7}) == (FORK_DEFAULT, FORK_BURN_IN)      # one chain of a larger job
    assert chain_fork(444, {"fork": 1}) == (0, 0)
    assert chain_fork(444, {"fork": 6, "fork.burn.in": 2}) == (6, 2)
    assert chain_fork(444, {"injected": object()}) == (0, 0)
    assert chain_burn_in(1, 444) == 100 and chain_burn_in(1, 1) == 1 and chain_burn_in(1, 444, {"burn.in.exact": True}) == 1
