"""GPU tests of tmvrnorm.ExpT (minimax tilting; TruncatedNormal is not vendored -> the checker is
the independent numpy restatement of the published algorithm in oracle/expt_oracle.py)."""
import numpy as np
import pytest
from scipy import stats

pytestmark = pytest.mark.gpu


def _cases():
    from lineqgpr_b200 import workloads as W
    rng = np.random.default_rng(4)
    B = rng.standard_normal((6, 6)); S = B @ B.T / 6 + 0.2 * np.eye(6); mu = 0.2 * rng.standard_normal(6)
    yield "random_d6", dict(mu=mu, Sigma=S, lb=mu - rng.uniform(0.2, 1.0, 6), ub=mu + rng.uniform(0.2, 1.5, 6))
    yield "tail_d3", dict(mu=np.zeros(3), Sigma=np.array([[1.0, 0.5, 0.2], [0.5, 1.0, 0.3], [0.2, 0.3, 1.0]]),
                          lb=np.array([1.5, -np.inf, 0.5]), ub=np.array([np.inf, -1.0, 0.9]))
    yield "cfg2", W.cfg2()[0]
    yield "cfg3", W.cfg3()[0]


@pytest.mark.parametrize("case", list(_cases()), ids=lambda c: c[0])
def test_expt_kernel_matches_oracle_draw_for_draw(lqg, case):
    """Same set-up, same Philox-addressed uniforms: same accepted proposals, Z to 1e-9."""
    from oracle import expt_oracle as EO
    from lineqgpr_b200 import expt
    name, par = case
    mu = np.asarray(par["mu"], float)
    su = expt.setup(par["lb"] - mu, par["ub"] - mu, par["Sigma"])
    so = EO.setup(par["lb"] - mu, par["ub"] - mu, par["Sigma"])
    assert np.array_equal(su["perm"], so["perm"])
    assert abs(su["psistar"] - so["psistar"]) < 1e-9 and np.abs(su["mu"] - so["mu"]).max() < 1e-7
    n, seed = 150, 11
    Zg = lqg.tmvrnorm(dict(par, **{"class": "ExpT"}), n, {"seed": seed, "setup": su, "return.Z": True})
    used = lqg.last_stats["proposals"]
    _, Zo, idx = EO.mvrandn(par["lb"] - mu, par["ub"] - mu, par["Sigma"], n, seed, su=su)
    assert used == idx[-1] + 1
    assert np.abs(Zg - Zo).max() < 1e-9 * max(1.0, np.abs(Zo).max())
    x = lqg.tmvrnorm(dict(par, **{"class": "ExpT"}), n, {"seed": seed, "setup": su})
    xo = (su["Lfull"] @ Zo)[su["order"], :] + mu[:, None]
    assert np.abs(x - xo).max() < 1e-9 * max(1.0, np.abs(xo).max())
    tol = 1e-9
    assert np.all(x >= np.asarray(par["lb"])[:, None] - tol) and np.all(x <= np.asarray(par["ub"])[:, None] + tol)


def test_expt_one_dimensional_is_truncated_normal(lqg):
    for (m, sd, lb, ub) in [(0.3, 1.7, -0.5, 1.0), (0.0, 1.0, 3.0, np.inf), (1.0, 0.5, -np.inf, -0.5)]:
        x = lqg.tmvrnorm(dict(mu=[m], Sigma=[[sd ** 2]], lb=[lb], ub=[ub]), 4000, {"seed": 3}, method="ExpT").ravel()
        assert x.min() >= lb and x.max() <= ub
        assert lqg.last_stats["proposals"] == 4000             # d = 1: every tilted proposal is accepted
        cdf = stats.truncnorm((lb - m) / sd, (ub - m) / sd, loc=m, scale=sd).cdf
        assert stats.kstest(x, cdf).pvalue > 0.01


def test_expt_agrees_with_rsm_and_hmc(lqg):
    """Three samplers, one target.  RSM (Maatouk & Bay) is exact when the TMVN mean is 0 and `map`
    is the mode of N(0, Sigma) on the box (the setting its ratio assumes, quirk Q1), so the
    ExpT/RSM comparison is made there; ExpT/HMC is compared on cfg2 as it comes."""
    from lineqgpr_b200 import workloads as W
    from lineqgpr_b200.model import solve_qp
    rng = np.random.default_rng(8)
    d = 6
    B = rng.standard_normal((d, d)); S = B @ B.T / d + 0.3 * np.eye(d)
    lb = np.array([0.3, -1.0, -np.inf, 0.1, -2.0, -0.5]); ub = np.array([2.0, 1.5, 0.8, np.inf, -0.2, 0.9])
    fin_l, fin_u = np.isfinite(lb), np.isfinite(ub)
    A = np.vstack([np.eye(d)[fin_l], -np.eye(d)[fin_u]]); b = np.concatenate([lb[fin_l], -ub[fin_u]])
    mode = solve_qp(np.linalg.inv(S), np.zeros(d), A, b)              # argmin x' S^-1 x / 2 on the box
    mode = np.minimum(np.maximum(mode, lb), ub)
    obj = dict(mu=np.zeros(d), map=mode, Sigma=S, lb=lb, ub=ub)
    n = 6000
    xr = lqg.tmvrnorm(obj, n, {"seed": 2}, method="RSM")
    xe = lqg.tmvrnorm(obj, n, {"seed": 3}, method="ExpT")
    p_rsm = np.array([stats.ks_2samp(xe[i], xr[i]).pvalue for i in range(d)])
    assert p_rsm.min() > 0.002, p_rsm
    par, _ = W.cfg2()
    n = 4000
    xe = lqg.tmvrnorm(dict(par, **{"class": "ExpT"}), n, {"seed": 1})
    rate = n / lqg.last_stats["proposals"]
    xh = lqg.tmvrnorm(dict(par, **{"class": "HMC"}), 4 * n, {"seed": 4, "burn.in": 50, "n.chains": 100})[:, ::4]
    p_hmc = np.array([stats.ks_2samp(xe[i], xh[i]).pvalue for i in range(20)])
    assert (p_hmc < 0.01).sum() <= 1, p_hmc.min()
    assert rate > 0.05


def test_simulate_with_the_package_default_sampler(lqg):
    """create() leaves sampler = "ExpT" (R/lineqGP.R:263): simulate must now work out of the box."""
    from lineqgpr_b200 import workloads as W
    par, aux = W.cfg1(m=30, sampler="ExpT")
    mdl = aux["lineq_model"]
    assert mdl["localParam"]["sampler"] == "ExpT"
    sim = lqg.simulate(mdl, nsim=500, seed=3, xtest=aux["xtest"])
    assert sim["xi.sim"].shape == (30, 500) and np.diff(sim["xi.sim"], axis=0).min() > -1e-9
    assert sim["stats"]["sampler"] == "ExpT"


@pytest.mark.parametrize("case", ["cfg1", "cfg2", "random60", "d1"])
def test_expt_setup_on_the_device_matches_the_host_restatement(lqg, case):
    """lqg_expt_setup (permuted Cholesky + Newton tilt on the GPU) against expt.setup: same permutation, factor,
    scaled bounds, tilt and psi* (1e-9), and tmvrnorm.ExpT with control["setup"] = "device" draws the same samples."""
    from lineqgpr_b200 import expt, workloads as W
    if case in ("cfg1", "cfg2"):
        par, _ = W.CONFIGS[case](sampler="ExpT")
        mu, S = np.asarray(par["mu"], float), np.asarray(par["Sigma"], float)
        l, u = np.asarray(par["lb"], float) - mu, np.asarray(par["ub"], float) - mu
    else:
        rng = np.random.default_rng(4)
        d = 60 if case == "random60" else 1
        B = rng.standard_normal((d, d)); S = B @ B.T / d + 0.3 * np.eye(d)
        mu = np.zeros(d); l = -0.5 - rng.uniform(size=d); u = 0.7 + rng.uniform(size=d)
        l[::7] = -np.inf
        par = dict(mu=mu, Sigma=S, lb=l, ub=u)
    h = expt.setup(l, u, S)
    g = expt.setup_device(l, u, S)
    d = l.size
    ph, pg = h["perm"], g["perm"]
    assert np.array_equal(np.sort(pg), np.arange(d))
    # the factor belongs to the permutation the device chose
    Sp = S[np.ix_(pg, pg)]
    assert np.abs(g["Lfull"] @ g["Lfull"].T - Sp).max() <= 1e-12 * np.abs(S).max() * d
    assert np.abs(g["L0"] - (g["Lfull"] / np.diag(g["Lfull"])[:, None] - np.eye(d))).max() <= 1e-13
    # Late in the order the remaining variables are all but unconstrained given the earlier ones: their
    # conditional log-masses are zero to rounding and the arg-min between them is decided by the last bit of
    # differently ordered sums, so the two permutations may part ways there (any order is a valid
    # factorization; it only matters for efficiency where the masses differ).  They must agree while the
    # pivot choice is decisive, and wherever they agree the factors agree.
    same = int(np.argmax(ph != pg)) if np.any(ph != pg) else d
    assert same >= min(d, 8)
    assert np.abs(h["Lfull"][:same, :same] - g["Lfull"][:same, :same]).max() <= 1e-11 * max(1.0, np.abs(h["Lfull"]).max())
    fin = np.isfinite(h["l"][:same])
    if fin.any():
        assert np.abs(h["l"][:same][fin] - g["l"][:same][fin]).max() <= 1e-10 * max(1.0, np.abs(h["l"][:same][fin]).max())
    # the tilt, compared variable by variable, and psi*
    mh = np.zeros(d); mg = np.zeros(d); mh[ph] = h["mu"]; mg[pg] = g["mu"]
    assert np.abs(mh - mg).max() <= 1e-8 * max(1.0, np.abs(mh).max())
    assert abs(h["psistar"] - g["psistar"]) <= 1e-9 * max(1.0, abs(h["psistar"]))
    obj = dict(par, **{"class": "ExpT"})
    xa = lqg.tmvrnorm(obj, 4000, {"seed": 3, "setup": g})
    sa = dict(lqg.last_stats)
    xb = lqg.tmvrnorm(obj, 4000, {"seed": 3, "setup": h})
    sb = dict(lqg.last_stats)
    lbv, ubv = np.asarray(par["lb"], float), np.asarray(par["ub"], float)
    assert np.all(xa >= lbv[:, None]) and np.all(xa <= ubv[:, None])
    se = np.sqrt((xa.var(1) + xb.var(1)) / 4000)
    assert (np.abs(xa.mean(1) - xb.mean(1)) > 4 * se + 1e-12).sum() <= max(1, d // 50)
    ra, rb = sa["accepted"] / sa["proposals"], sb["accepted"] / sb["proposals"]
    assert abs(ra - rb) <= 0.05 * max(ra, rb) + 0.01                  # same acceptance rate: the same psi*
    if same == d:
        assert np.abs(xa - xb).max() <= 1e-8 * max(1.0, np.abs(xb).max())
    xc = lqg.tmvrnorm(obj, 64, {"seed": 3, "setup": "device"})
    assert np.array_equal(xa[:, :64], lqg.tmvrnorm(obj, 64, {"seed": 3, "setup": g})) and np.array_equal(xc, xa[:, :64])
