"""GPU parity tests of the HMC path, all through the C ABI (ctypes), against the oracle and the
committed golden vectors (outputs of the compiled reference source).

Tolerance: north_star asks for trajectories within 1e-10 relative (FP64) under injected Gaussian
draws.  RTOL below is that bound, relative to the largest coordinate of the compared state."""
import numpy as np
import pytest
from scipy import stats

pytestmark = pytest.mark.gpu
RTOL = 1e-10

SMALL = ["rd_example_2d", "dense_d10_c8", "roxygen_box_d100", "cfg1_canonical"]


def rel_err(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("name", SMALL)
def test_canonical_chain_matches_reference_golden(lqg, golden, name, mode):
    """Free-running chain, same injected draws as the compiled reference consumed; both the
    per-chain kernel (mode 1) and the lock-step GEMM rounds (mode 2)."""
    G = golden(name)
    n = int(G["n"])
    z, st = lqg.rtmg_canonical(n, 0, G["z0"], G["F"], G["g"], injected=G["draws"], mode=mode)
    assert rel_err(z.T, G["z_ref"]) < RTOL
    assert st["bounces"] == int(G["bounces"])
    assert st["vel_restarts"] == int(G["vel_restarts"]) and st["chk_restarts"] == int(G["chk_restarts"])
    assert st["violations"] == 0 and st["transitions"] == n


@pytest.mark.parametrize("name", SMALL + ["stiff_d60_c200"])
def test_canonical_teacher_forced_transitions(lqg, oracle, golden, name):
    """Per-transition parity: K chains, chain k starts at the reference state k-1 and gets the
    draws of a fresh stream; one transition each; compared with the oracle's transition."""
    G = golden(name)
    d = G["z0"].size
    states = np.vstack([G["z0"][None, :], G["z_ref"][:-1]])
    K = min(states.shape[0], 24)
    L = 400 * d
    rng = np.random.default_rng(17)
    draws = rng.standard_normal((K, L))
    want = np.zeros((K, d)); nb = np.zeros(K, dtype=int); nr = 0
    for k in range(K):
        zk, info = oracle.rtmg_canonical(1, 0, states[k], G["F"], G["g"], injected=draws[k])
        assert info["rc"] == 0
        want[k] = zk[0]; nb[k] = info["bounces"]; nr += info["vel_restarts"] + info["chk_restarts"]
    z, st = lqg.rtmg_canonical(1, 0, states[:K].T, G["F"], G["g"], n_chains=K, injected=draws)
    assert (G["F"] @ z + G["g"][:, None]).min() >= -1e-12 and st["violations"] == 0
    # Billiard dynamics amplify rounding exponentially in the number of bounces (SURVEY hard part 2):
    # the 1e-10 bound is asserted for transitions of <= 64 bounces; longer ones (only the stiff
    # fixture has them, ~1900 bounces per transition) are covered by constraint satisfaction above.
    short = nb <= 64
    if name != "stiff_d60_c200":
        assert short.all()
        assert st["bounces"] == nb.sum() and st["vel_restarts"] + st["chk_restarts"] == nr
    if short.any():
        assert rel_err(z.T[short], want[short]) < RTOL
    else:
        assert abs(st["bounces"] - nb.sum()) < 0.5 * nb.sum()     # same regime, different micro-trajectory


@pytest.mark.parametrize("name", SMALL + ["stiff_d60_c200"])
def test_hit_search_teacher_forced_per_bounce(lqg, oracle, golden, name):
    """Event-level parity: for K (state, momentum) pairs the sequence of wall hits (t, cn) of the
    GPU chain is compared with the oracle's trace.  The FIRST event is a pure teacher-forced
    _getNextLinearHitTime (same a, b in); later events are compared until rounding, amplified by
    the billiard dynamics, first changes the hit wall — which the stiff fixture needs many
    bounces to do, so its first events still pin ties, dead zones and the arg-min."""
    G = golden(name)
    d = G["z0"].size
    states = np.vstack([G["z0"][None, :], G["z_ref"][:-1]])
    K = min(states.shape[0], 12)
    cap = 64
    rng = np.random.default_rng(23)
    draws = rng.standard_normal((K, 60 * d))
    z, tt, tc, tn = lqg.rtmg_canonical_trace(1, states[:K].T, G["F"], G["g"], draws, n_chains=K, trace_cap=cap)
    matched = []
    for k in range(K):
        _, info = oracle.rtmg_canonical(1, 0, states[k], G["F"], G["g"], injected=draws[k], trace_cap=cap)
        n = min(info["n_trace"], tn[k], cap)
        m = 0
        while m < n and tc[k, m] == info["trace_cn"][m] and abs(tt[k, m] - info["trace_t"][m]) <= 1e-10 * max(1.0, info["trace_t"][m]) * 4.0 ** min(m, 12):
            m += 1
        if n > 0:                                   # the first hit search is exactly teacher forced
            assert tc[k, 0] == info["trace_cn"][0], (name, k)
            assert abs(tt[k, 0] - info["trace_t"][0]) <= 1e-12 * max(1.0, info["trace_t"][0]), (name, k)
        matched.append((m, n))
    if name != "stiff_d60_c200":
        assert all(m == n for m, n in matched), matched       # whole transitions agree event by event
    else:
        assert np.mean([m for m, n in matched]) >= 8, matched  # chaotic: long agreeing prefixes


def test_canonical_constraints_and_layout(lqg, golden):
    G = golden("dense_d10_c8")
    z, st = lqg.rtmg_canonical(40, 5, G["z0"], G["F"], G["g"], n_chains=7, burn_in=3)
    assert z.shape == (10, 7 * 40) and st["transitions"] == 7 * 43
    assert (G["F"] @ z + G["g"][:, None]).min() >= -1e-12
    assert st["violations"] == 0
    # Philox streams are keyed by chain: chain 0 of a 7-chain run == a 1-chain run
    z1, _ = lqg.rtmg_canonical(40, 5, G["z0"], G["F"], G["g"], n_chains=1, burn_in=3)
    assert np.array_equal(z1, z[:, :40])


def test_lqg_rtmg_entry_point_mirrors_dot_call(lqg, golden):
    """lqg_rtmg(n, seed, initial, numlin, F, g, numquad, quads) -> n x d, row = sample."""
    from lineqgpr_b200 import _lib as L
    G = golden("rd_example_2d")
    n, d = 50, 2
    F = L.f64(G["F"]); g = np.ascontiguousarray(G["g"]); z0 = np.ascontiguousarray(G["z0"])
    out = np.zeros((n, d), order="F")
    rc = L.lib().lqg_rtmg(n, 123, L.ptr(z0), d, 2, L.ptr(F), L.ptr(g), 0, None, L.ptr(out))
    assert rc == 0
    assert (G["F"] @ out.T + G["g"][:, None]).min() >= 0
    z, _ = lqg.rtmg_canonical(n, 123, G["z0"], G["F"], G["g"])
    assert np.array_equal(out, z.T)
    rc = L.lib().lqg_rtmg(n, 123, L.ptr(z0), d, 2, L.ptr(F), L.ptr(g), 1, None, L.ptr(out))
    assert rc == 2 and "quadratic" in L.last_error()


# ------------------------------------------------------------------------------------- box form
def _box_problem(oracle, par):
    H, r, init, f, g = oracle.hmc_setup(par["mu"], par["Sigma"], par["lb"], par["ub"], par.get("map"))
    w = oracle.whiten(H, r, init, f, g)
    return w, init


def _box_cases():
    from lineqgpr_b200 import workloads as W
    rng = np.random.default_rng(3)
    B = rng.standard_normal((12, 12)); S = B @ B.T / 12 + 0.05 * np.eye(12); mu = 0.3 * rng.standard_normal(12)
    lb = np.where(rng.uniform(size=12) < 0.3, -np.inf, mu - rng.uniform(0.05, 1.0, 12))
    ub = np.where(rng.uniform(size=12) < 0.3, np.inf, mu + rng.uniform(0.05, 1.0, 12))
    yield "random_d12", dict(mu=mu, Sigma=S, lb=lb, ub=ub), 60
    yield "cfg2", W.cfg2()[0], 80
    yield "cfg3", W.cfg3()[0], 80
    yield "cfg1", W.cfg1()[0], 30
    yield "cfg4_m100", W.cfg4(m=100)[0], 12        # d = 200 stacked bounded + monotone, 2 warps/chain
    yield "cfg5_D30", W.cfg5(D=30, knots=10, n=150, n_test=10)[0], 12   # d = 300, 4 warps/chain


@pytest.mark.parametrize("case", list(_box_cases()), ids=lambda c: c[0])
def test_box_chain_matches_oracle_canonical_chain(lqg, oracle, case):
    """lqg_hmc_box (eta frame) vs the oracle's whitened chain un-whitened (tmg:R/tmg.R:106),
    same U = Ri and same injected draws; free-running short horizon + teacher-forced steps."""
    name, par, n = case
    w, init = _box_problem(oracle, par)
    d = init.size
    seed = 77
    z, info = oracle.rtmg_canonical(n, seed, w["initial2"], w["f2"], w["g2"])
    x_ref = z @ w["Ri"].T + w["Mir"][None, :]
    draws = oracle.normal_stream(seed, info["draws"])
    ctl = {"burn.in": 0, "n.chains": 1, "U": w["Ri"], "injected": draws[None, :]}
    x = lqg.tmvrnorm(dict(par, **{"class": "HMC"}), n, ctl)
    st = dict(lqg.last_stats)
    assert st["violations"] == 0
    lbv, ubv = np.asarray(par["lb"]), np.asarray(par["ub"])
    assert np.all(x >= lbv[:, None]) and np.all(x <= ubv[:, None])            # exact, no tolerance
    # teacher-forced: K chains started from consecutive reference states
    K = min(n, 16)
    rng = np.random.default_rng(5)
    dr = rng.standard_normal((K, 300 * d))
    starts_z = np.vstack([w["initial2"][None, :], z[:-1]])[:K]
    want = np.zeros((K, d))
    for k in range(K):
        zk, ik = oracle.rtmg_canonical(1, 0, starts_z[k], w["f2"], w["g2"], injected=dr[k])
        want[k] = zk[0] @ w["Ri"].T + w["Mir"]
    starts_x = starts_z @ w["Ri"].T + w["Mir"][None, :]
    # the start must be strictly inside the box for the C ABI's feasibility check
    starts_x = np.minimum(np.maximum(starts_x, lbv + 1e-13), ubv - 1e-13)
    from lineqgpr_b200 import _lib as L
    out = np.empty((d, K), order="F")
    hs = L.HmcStats()
    opts = L.make_opts(n_chains=K, injected=np.ascontiguousarray(dr), n_injected=dr.shape[1], prepass=0)
    # The whitened chain samples N(Mir, Ri Ri') — the centre and covariance as tmg's whitening left them
    # (tmg:R/tmg.R:15-28,51), which differ from (mu, Sigma) by ~cond(Sigma) eps.  Given THOSE inputs the
    # eta-frame kernel must reproduce the oracle's transition to north_star's 1e-10 ...
    Uf, Sf = L.f64(w["Ri"]), L.f64(w["Ri"] @ w["Ri"].T)
    mu_c = np.ascontiguousarray(w["Mir"], dtype=np.float64)
    lb_c = np.ascontiguousarray(lbv, dtype=np.float64); ub_c = np.ascontiguousarray(ubv, dtype=np.float64)
    sx = np.ascontiguousarray(starts_x)
    rc = L.lib().lqg_hmc_box(d, L.ptr(mu_c), L.ptr(Uf), 1, L.ptr(Sf), L.ptr(lb_c), L.ptr(ub_c), L.ptr(sx), d,
                             1, 0, opts, L.ptr(out), None, hs)
    L.check(rc, "lqg_hmc_box")
    scale = np.abs(want).max()
    assert np.abs(out.T - want).max() / scale < RTOL, name
    # ... and with the caller's own (mu, Sigma) it stays within the whitening residual of it
    Sf2 = L.f64(par["Sigma"]); mu2 = np.ascontiguousarray(par["mu"], dtype=np.float64)
    rc = L.lib().lqg_hmc_box(d, L.ptr(mu2), L.ptr(Uf), 1, L.ptr(Sf2), L.ptr(lb_c), L.ptr(ub_c), L.ptr(sx), d,
                             1, 0, opts, L.ptr(out), None, hs)
    L.check(rc, "lqg_hmc_box")
    assert np.abs(out.T - want).max() / scale < 1e-8, name
    # free-running: agree until rounding flips a discrete event; the first samples must match
    head = max(2, min(n, 5))
    assert np.abs(x[:, :head].T - x_ref[:head]).max() / np.abs(x_ref).max() < 1e-8, name


def test_box_queue_overflow_is_bitwise_neutral(lqg):
    """A tight box (+-0.3 sd, d = 128) gives ~85 surviving walls per search in ONE warp: the
    64-entry candidate queue overflows and is flushed mid-search all the time.  Padding the same
    problem with unconstrained, independent coordinates (d = 300) maps it to the 4-warp kernel
    variant, where no warp overflows, without changing a single operation on the constrained
    block — so the two runs must agree BIT FOR BIT (chaos-proof, unlike an oracle comparison:
    one transition here is ~1900 bounces)."""
    d, pad = 128, 172
    rng = np.random.default_rng(d)
    B = rng.standard_normal((d, d)); S = B @ B.T / d + 0.5 * np.eye(d)
    mu = 0.1 * rng.standard_normal(d); sd = np.sqrt(np.diag(S))
    lb, ub = mu - 0.3 * sd, mu + 0.3 * sd
    from lineqgpr_b200.samplers import whiten_box
    U = whiten_box(mu, S)
    n, chains = 2, 3
    draws = rng.standard_normal((chains, 400, d))                        # 400 momenta per chain (restarts are frequent here)
    ctl = {"burn.in": 0, "n.chains": chains, "max.restarts": 1000, "hmc.search": 2}   # 2: no time windows, every search sees ~85 walls
    x1 = lqg.tmvrnorm(dict(mu=mu, Sigma=S, lb=lb, ub=ub), n * chains, dict(ctl, U=U, injected=draws.reshape(chains, -1)), method="HMC")
    s1 = dict(lqg.last_stats)
    S2 = np.eye(d + pad); S2[:d, :d] = S
    U2 = np.eye(d + pad); U2[:d, :d] = U
    mu2 = np.concatenate([mu, np.zeros(pad)])
    lb2 = np.concatenate([lb, np.full(pad, -np.inf)]); ub2 = np.concatenate([ub, np.full(pad, np.inf)])
    draws2 = np.concatenate([draws, rng.standard_normal((chains, 400, pad))], axis=2)
    x2 = lqg.tmvrnorm(dict(mu=mu2, Sigma=S2, lb=lb2, ub=ub2), n * chains, dict(ctl, U=U2, injected=draws2.reshape(chains, -1)), method="HMC")
    s2 = dict(lqg.last_stats)
    assert s1["exact_evals"] / s1["hit_searches"] > 70                   # single warp: more than one queue-full
    assert s1["bounces"] == s2["bounces"] and s1["vel_restarts"] == s2["vel_restarts"] and s1["chk_restarts"] == s2["chk_restarts"]
    assert np.array_equal(x1, x2[:d])
    assert s1["violations"] == 0 and np.all(x1 >= lb[:, None]) and np.all(x1 <= ub[:, None])


def test_box_errors(lqg):
    from lineqgpr_b200 import _lib as L
    obj = dict(mu=np.zeros(3), Sigma=np.eye(3), lb=-np.ones(3), ub=np.ones(3), map=np.array([0.0, 5.0, 0.0]))
    # clamp keeps the start inside: fine
    x = lqg.tmvrnorm(obj, 10, {"burn.in": 2}, method="HMC")
    assert x.shape == (3, 10)
    mu = np.zeros(3); S = L.f64(np.eye(3)); lb = -np.ones(3); ub = np.ones(3); bad = np.array([0.0, 1.0, 0.0])
    out = np.zeros((3, 4), order="F")
    rc = L.lib().lqg_hmc_box(3, L.ptr(mu), L.ptr(S), 1, L.ptr(S), L.ptr(lb), L.ptr(ub), L.ptr(bad), 0, 4, 0,
                             L.make_opts(), L.ptr(out), None, None)
    assert rc == 3 and "initial point violates" in L.last_error()
    rc = L.lib().lqg_hmc_box(3, L.ptr(mu), L.ptr(S), 1, L.ptr(S), L.ptr(ub), L.ptr(lb), L.ptr(mu), 0, 4, 0,
                             L.make_opts(), L.ptr(out), None, None)
    assert rc == 1


def test_box_unconstrained_limit_is_gaussian(lqg):
    """Anchor: bounds far away -> each transition returns mu + U a, i.i.d. N(mu, Sigma)."""
    rng = np.random.default_rng(0)
    B = rng.standard_normal((5, 5)); S = B @ B.T + 0.5 * np.eye(5); mu = rng.standard_normal(5)
    obj = dict(mu=mu, Sigma=S, lb=np.full(5, -1e6), ub=np.full(5, np.inf))
    x = lqg.tmvrnorm(obj, 20000, {"burn.in": 0, "n.chains": 64, "seed": 3}, method="HMC")
    assert lqg.last_stats["bounces"] == 0
    assert np.abs(x.mean(1) - mu).max() < 4 * np.sqrt(np.diag(S).max() / 20000)
    assert np.abs(np.cov(x) - S).max() < 0.08 * np.abs(S).max()
    for i in range(5):
        assert stats.kstest((x[i] - mu[i]) / np.sqrt(S[i, i]), "norm").pvalue > 1e-3


def test_box_prepass_equals_in_kernel_draws(lqg):
    """The DMMA pre-pass and the per-chain product draw the same Philox normals."""
    from lineqgpr_b200 import workloads as W
    par, _ = W.cfg5(D=30, knots=10, n=150, n_test=10)          # d = 300
    a = lqg.tmvrnorm(dict(par, **{"class": "HMC"}), 64, {"burn.in": 2, "n.chains": 32, "seed": 11, "prepass": 0})
    sa = dict(lqg.last_stats)
    b = lqg.tmvrnorm(dict(par, **{"class": "HMC"}), 64, {"burn.in": 2, "n.chains": 32, "seed": 11, "prepass": 1})
    sb = dict(lqg.last_stats)
    assert sb["prepass_ms"] > 0 and sa["prepass_ms"] == 0
    # same draws, different summation order in U a: first emitted sample of every chain agrees tightly
    assert np.abs(a[:, ::2] - b[:, ::2]).max() / np.abs(a).max() < 1e-7
    assert sa["violations"] == 0 and sb["violations"] == 0


def test_box_sharding_invariance(lqg):
    """chain_offset: 8 chains in one call == two calls of 4 chains (what the multi-GPU launcher does)."""
    from lineqgpr_b200 import workloads as W
    par, _ = W.cfg3()
    obj = dict(par, **{"class": "HMC"})
    full = lqg.tmvrnorm(obj, 80, {"burn.in": 5, "n.chains": 8, "seed": 9})
    lo = lqg.tmvrnorm(obj, 40, {"burn.in": 5, "n.chains": 4, "seed": 9, "chain.offset": 0})
    hi = lqg.tmvrnorm(obj, 40, {"burn.in": 5, "n.chains": 4, "seed": 9, "chain.offset": 4})
    assert np.array_equal(full, np.hstack([lo, hi]))


def test_forked_families_share_a_root_chain(lqg):
    """lqg_opts.fork = F: the chains of a family start from the state their root chain reached after the
    burn-in.  (1) With fork.burn.in = 0 the first chain of every family IS the independent chain of the
    same global id, bit for bit (it continues the root's stream); the other chains of the family differ
    from their independent versions but start from the same point.  (2) Which root a chain has follows
    from its global id: shards that cut through families give the columns of the whole job.
    (3) Statistics of the forked job agree with independent chains."""
    from lineqgpr_b200 import workloads as W
    par, _ = W.cfg3()
    obj = dict(par, **{"class": "HMC"})
    n_ch, n_per = 12, 6
    indep = lqg.tmvrnorm(obj, n_ch * n_per, {"burn.in": 7, "n.chains": n_ch, "seed": 3, "fork": 1, "burn.in.exact": True})
    s_ind = dict(lqg.last_stats)
    fork0 = lqg.tmvrnorm(obj, n_ch * n_per, {"burn.in": 7, "n.chains": n_ch, "seed": 3, "fork": 3, "fork.burn.in": 0, "burn.in.exact": True})
    s_f = dict(lqg.last_stats)
    for c in range(n_ch):
        a, b = indep[:, c * n_per:(c + 1) * n_per], fork0[:, c * n_per:(c + 1) * n_per]
        assert np.array_equal(a, b) == (c % 3 == 0), c
    assert s_f["transitions"] == 4 * 7 + n_ch * n_per and s_ind["transitions"] == n_ch * (7 + n_per)
    # (2) shards of 5 + 7 chains cut through the families {3,4,5} and nothing else lines up with 3
    ctl = {"burn.in": 7, "seed": 3, "fork": 3, "fork.burn.in": 2, "burn.in.exact": True}
    whole = lqg.tmvrnorm(obj, n_ch * n_per, dict(ctl, **{"n.chains": n_ch}))
    lo = lqg.tmvrnorm(obj, 5 * n_per, dict(ctl, **{"n.chains": 5, "chain.offset": 0}))
    hi = lqg.tmvrnorm(obj, 7 * n_per, dict(ctl, **{"n.chains": 7, "chain.offset": 5}))
    assert np.array_equal(whole, np.hstack([lo, hi]))
    one = lqg.tmvrnorm(obj, n_per, dict(ctl, **{"n.chains": 1, "chain.offset": 4}))
    assert np.array_equal(one, whole[:, 4 * n_per:5 * n_per])
    # (3) moments: 300 chains in families (the wrapper's default) against 300 independent chains
    x3 = lqg.tmvrnorm(obj, 6000, {"burn.in": 100, "n.chains": 300, "seed": 5})
    assert lqg.last_stats["fork"] == 6
    x1 = lqg.tmvrnorm(obj, 6000, {"burn.in": 100, "n.chains": 300, "seed": 6, "fork": 1})
    se = np.sqrt((x3.var(1) + x1.var(1)) / 6000)
    assert (np.abs(x3.mean(1) - x1.mean(1)) > 3.5 * se).sum() <= 1
    pv = np.array([stats.ks_2samp(x3[i], x1[i]).pvalue for i in range(x3.shape[0])])
    assert (pv < 0.01).sum() <= max(1, int(0.05 * x3.shape[0]))


@pytest.mark.parametrize("cfg", ["cfg1", "cfg2", "cfg3"])
def test_box_statistical_agreement_with_oracle(lqg, oracle, cfg):
    """L-stat: per-coordinate KS p > 0.01 and mean/variance within 3 Monte-Carlo standard errors
    of an oracle chain (libstdc++ RNG) on the same (mu, Sigma, lb, ub)."""
    from lineqgpr_b200 import workloads as W
    par, _ = W.CONFIGS[cfg]()
    n = 6000
    xo, _ = oracle.tmvrnorm_HMC(par, n, {"burn.in": 100}, seed=4)
    xg = lqg.tmvrnorm(dict(par, **{"class": "HMC"}), n, {"burn.in": 100, "n.chains": 60, "seed": 8})
    assert lqg.last_stats["violations"] == 0
    thin = 4
    xo_t, xg_t = xo[:, ::thin], xg[:, ::thin]
    d = xo.shape[0]
    pvals = np.array([stats.ks_2samp(xo_t[i], xg_t[i]).pvalue for i in range(d)])
    # d simultaneous tests at level 0.01.  Neighbouring coordinates of a GP are correlated at 0.9 and more, so
    # chance rejections come in runs of 5-15 coordinates (one unlucky stretch of either chain): the raw count
    # gets room for one such run, and the same test on decorrelated coordinates (projections on the
    # eigenvectors of Sigma, nearly independent tests) keeps the 5 % allowance.
    assert (pvals < 0.01).sum() <= max(1, int(0.15 * d)), ((pvals < 0.01).sum(), pvals.min())
    V = np.linalg.eigh(np.asarray(par["Sigma"], float))[1]
    po, pg = V.T @ xo_t, V.T @ xg_t
    pv_dec = np.array([stats.ks_2samp(po[i], pg[i]).pvalue for i in range(d)])
    assert (pv_dec < 0.01).sum() <= max(1, int(0.05 * d)), ((pv_dec < 0.01).sum(), pv_dec.min())
    ne = xo_t.shape[1]
    se_mean = np.sqrt((xo_t.var(1) + xg_t.var(1)) / ne)
    assert (np.abs(xo_t.mean(1) - xg_t.mean(1)) > 3 * se_mean).sum() <= max(1, int(0.05 * d))
    se_var = np.sqrt(2.0 / (ne - 1)) * np.sqrt(xo_t.var(1) ** 2 + xg_t.var(1) ** 2) * 1.5   # kurtosis slack
    assert (np.abs(xo_t.var(1) - xg_t.var(1)) > 3 * se_var).sum() <= max(1, int(0.05 * d))


def test_multi_chain_default_with_burn_in_one_matches_a_long_single_chain(lqg):
    """simulate.* hands tmvrnorm burn.in = 1 (R/lineqGP.R:617): fine for the reference's ONE long chain,
    but with many chains started at the same clamped MAP every column would sit two transitions from it.
    The wrapper therefore runs at least 100 burn-in transitions per chain (chain_burn_in); on a tightly
    constrained problem (MAP on the boundary of a narrow box) the default multi-chain output must have the
    distribution of a single long chain.  (With T = pi/2 the exact HMC forgets its start within a few
    transitions — lag-1 autocorrelations are below 0.1 at every BASELINE config — so the floor is a
    safeguard, not something a KS test on a few hundred columns can see.)"""
    rng = np.random.default_rng(5)
    d = 24
    B = rng.standard_normal((d, d)); S = B @ B.T / d + 0.2 * np.eye(d)
    sd = np.sqrt(np.diag(S))
    mu = np.zeros(d)
    lb, ub = 0.25 * sd, 2.0 * sd                                    # the box sits in the upper tail: the mode is in its lb corner
    par = dict(mu=mu, Sigma=S, lb=lb, ub=ub, map=lb + 0.02 * sd)
    par["class"] = "HMC"
    n = 4096
    one = lqg.tmvrnorm(par, 4 * n, {"burn.in": 200, "n.chains": 1, "seed": 3})[:, ::4]
    many = lqg.tmvrnorm(par, n, {"burn.in": 1, "n.chains": 512, "seed": 9})
    assert lqg.last_stats["burn_in_per_chain"] == 100
    pv = np.array([stats.ks_2samp(one[i], many[i]).pvalue for i in range(d)])
    assert (pv < 0.01).sum() <= 1, pv.min()
    se = np.sqrt((one.var(1) + many.var(1)) / n)
    assert (np.abs(one.mean(1) - many.mean(1)) > 3.5 * se).sum() <= 1
    lqg.tmvrnorm(par, 512, {"burn.in": 1, "n.chains": 512, "seed": 9, "burn.in.exact": True})
    assert lqg.last_stats["burn_in_per_chain"] == 1                 # the caller's value stands when asked for


def test_stacked_constraints_statistics_m150(lqg, oracle):
    """Bounded + monotone stacked system (the cfg4 structure, Lambda = [I; D], rank-deficient
    Sigma_eta + nugget) at m = 150 (d = 300, 4 warps per chain): moments vs an oracle chain."""
    from lineqgpr_b200 import workloads as W
    par, _ = W.cfg4(m=150)
    n = 3000
    xo, _ = oracle.tmvrnorm_HMC(par, n, {"burn.in": 200}, seed=12)
    xg = lqg.tmvrnorm(dict(par, **{"class": "HMC"}), n, {"burn.in": 200, "n.chains": 30, "seed": 21})
    assert lqg.last_stats["violations"] == 0
    thin = 5
    xo_t, xg_t = xo[:, ::thin], xg[:, ::thin]
    d = xo.shape[0]
    pvals = np.array([stats.ks_2samp(xo_t[i], xg_t[i]).pvalue for i in range(d)])
    assert (pvals < 0.01).sum() <= int(0.05 * d), ((pvals < 0.01).sum(), pvals.min())
    ne = xo_t.shape[1]
    se = np.sqrt((xo_t.var(1) + xg_t.var(1)) / ne)
    assert (np.abs(xo_t.mean(1) - xg_t.mean(1)) > 3 * se).sum() <= int(0.05 * d)


@pytest.mark.parametrize("cfg", ["cfg4", "cfg5"])
def test_full_size_teacher_forced_parity(lqg, oracle, cfg):
    """Trajectory parity AT the BASELINE shapes (cfg4: d = 2000, c = 2999, the 8-warp kernel
    variant; cfg5: d = 1000, c = 900, the 4-warp variant): K chains started at states of an oracle
    chain, one transition each with injected draws, compared with the oracle's transition."""
    from lineqgpr_b200 import workloads as W
    from lineqgpr_b200 import _lib as L
    par = W.cfg4()[0] if cfg == "cfg4" else W.cfg5(n_test=10)[0]
    H, r, init, f, g = oracle.hmc_setup(par["mu"], par["Sigma"], par["lb"], par["ub"], par.get("map"))
    w = oracle.whiten(H, r, init, f, g)
    d = init.size
    K = 6
    z, info = oracle.rtmg_canonical(K - 1, 3, w["initial2"], w["f2"], w["g2"])      # a short oracle chain for the starts
    starts_z = np.vstack([w["initial2"][None, :], z])
    rng = np.random.default_rng(9)
    dr = rng.standard_normal((K, 40 * d))
    want = np.zeros((K, d)); nb = np.zeros(K, int)
    for k in range(K):
        zk, ik = oracle.rtmg_canonical(1, 0, starts_z[k], w["f2"], w["g2"], injected=dr[k])
        assert ik["rc"] == 0
        want[k] = zk[0] @ w["Ri"].T + w["Mir"]; nb[k] = ik["bounces"]
    lbv, ubv = np.asarray(par["lb"], float), np.asarray(par["ub"], float)
    starts_x = starts_z @ w["Ri"].T + w["Mir"][None, :]
    starts_x = np.ascontiguousarray(np.minimum(np.maximum(starts_x, lbv + 1e-13), ubv - 1e-13))
    out = np.empty((d, K), order="F")
    hs = L.HmcStats()
    opts = L.make_opts(n_chains=K, injected=np.ascontiguousarray(dr), n_injected=dr.shape[1], prepass=0)
    # the inputs the whitened chain actually samples from: centre Mir and covariance Ri Ri' as tmg's
    # whitening produced them (see test_gpu_parity_full.py::test_full_size_transition_against_extended_precision)
    Uf, Sf = L.f64(w["Ri"]), L.f64(w["Ri"] @ w["Ri"].T)
    mu_c = np.ascontiguousarray(w["Mir"], dtype=np.float64)
    rc = L.lib().lqg_hmc_box(d, L.ptr(mu_c), L.ptr(Uf), 1, L.ptr(Sf), L.ptr(lbv), L.ptr(ubv), L.ptr(starts_x), d,
                             1, 0, opts, L.ptr(out), None, hs)
    L.check(rc, "lqg_hmc_box")
    assert hs.violations == 0
    assert np.all(out >= lbv[:, None]) and np.all(out <= ubv[:, None])
    err = np.abs(out.T - want).max(axis=1) / np.abs(want).max()
    assert hs.bounces == nb.sum(), (hs.bounces, nb.sum())
    assert err.max() < RTOL, err                              # north_star's 1e-10 at the benchmarked shape
    # the caller's own (mu, Sigma): off by the whitening residual cond(Sigma_eta) eps (~1e-9 at cfg4)
    Sf2 = L.f64(par["Sigma"]); mu2 = np.ascontiguousarray(par["mu"], dtype=np.float64)
    rc = L.lib().lqg_hmc_box(d, L.ptr(mu2), L.ptr(Uf), 1, L.ptr(Sf2), L.ptr(lbv), L.ptr(ubv), L.ptr(starts_x), d,
                             1, 0, opts, L.ptr(out), None, hs)
    L.check(rc, "lqg_hmc_box")
    err2 = np.abs(out.T - want).max(axis=1) / np.abs(want).max()
    assert hs.bounces == nb.sum() and err2.max() < 1e-7, err2


def test_full_size_m1000_properties(lqg):
    """BASELINE size (cfg4: d = 2000, c = 2999; cfg5: d = 1000, c = 900): size-independent
    properties — every emitted value inside the box, counters consistent, finite output."""
    from lineqgpr_b200 import workloads as W
    for par, n_chains, n in ((W.cfg4()[0], 296, 592), (W.cfg5(n_test=10)[0], 296, 592)):
        x = lqg.tmvrnorm(dict(par, **{"class": "HMC"}), n, {"burn.in": 3, "n.chains": n_chains, "seed": 1, "burn.in.exact": True})
        st = dict(lqg.last_stats)
        lb, ub = np.asarray(par["lb"]), np.asarray(par["ub"])
        assert np.isfinite(x).all()
        assert np.all(x >= lb[:, None]) and np.all(x <= ub[:, None])
        assert st["violations"] == 0 and st["failed_chains"] == 0
        # families of 6 chains share a root chain's 3 burn-in transitions, then every chain runs min(5, 3) more
        assert st["fork"] == 6 and st["transitions"] == -(-n_chains // 6) * 3 + n_chains * (3 + n // n_chains)
        assert st["hit_searches"] >= st["bounces"] + st["transitions"]


@pytest.mark.parametrize("d", [2500, 5000])
def test_kernel_shapes_beyond_2048_coordinates(lqg, d):
    """d in (2048, 4096] and (4096, 8192] run other shapes of the box kernel (8 and 16 warps per chain, 16
    coordinates per thread, register or recomputed thresholds): every emitted value inside the box, the
    three hit-search modes (tan(t/2) ranking with windows, without, tmg's arithmetic throughout) decide
    the same events — identical with and without windows, 1e-9 against the exact search."""
    rng = np.random.default_rng(1)
    B = rng.standard_normal((d, 64)); S = B @ B.T / 64 + np.eye(d)
    sd = np.sqrt(np.diag(S))
    par = dict(mu=np.zeros(d), Sigma=S, lb=-2.2 * sd, ub=2.6 * sd)
    par["class"] = "HMC"
    out = {}
    for mode in (0, 2, 1):
        out[mode] = lqg.tmvrnorm(par, 16, {"burn.in": 3, "n.chains": 8, "seed": 5, "burn.in.exact": True, "hmc.search": mode})
        st = dict(lqg.last_stats)
        assert st["violations"] == 0 and st["failed_chains"] == 0 and st["bounces"] > 10 * st["transitions"]
        assert np.all(out[mode] >= par["lb"][:, None]) and np.all(out[mode] <= par["ub"][:, None])
    assert np.array_equal(out[0], out[2])
    first = [out[m].reshape(d, 8, 2, order="F")[:, :, 0] for m in (2, 1)]
    assert np.abs(first[0] - first[1]).max() <= 1e-9 * np.abs(first[1]).max()


# ------------------------------------------------------------------------------------- edge cases
def test_edge_cases_shapes_and_bounds(lqg):
    """d = 1, one-sided and unbounded coordinates, more chains than samples, a single long chain."""
    # d = 1 box: exact truncated normal support, KS against scipy
    obj = dict(mu=[0.3], Sigma=[[1.7 ** 2]], lb=[-0.5], ub=[1.0])
    x = lqg.tmvrnorm(obj, 4000, {"burn.in": 20, "n.chains": 50, "seed": 2}, method="HMC").ravel()
    assert x.min() >= -0.5 and x.max() <= 1.0
    cdf = stats.truncnorm((-0.5 - 0.3) / 1.7, (1.0 - 0.3) / 1.7, loc=0.3, scale=1.7).cdf
    assert stats.kstest(x[::2], cdf).pvalue > 0.01
    # mixed: coordinate 0 unbounded, 1 lower only, 2 upper only
    rng = np.random.default_rng(1)
    B = rng.standard_normal((3, 3)); S = B @ B.T + 0.3 * np.eye(3)
    obj = dict(mu=np.zeros(3), Sigma=S, lb=np.array([-np.inf, 0.1, -np.inf]), ub=np.array([np.inf, np.inf, -0.2]),
               map=np.array([0.0, 0.5, -0.5]))
    x = lqg.tmvrnorm(obj, 2000, {"burn.in": 10, "seed": 3}, method="HMC")
    assert x.shape == (3, 2000) and x[1].min() >= 0.1 and x[2].max() <= -0.2
    assert lqg.last_stats["violations"] == 0
    # nsim smaller than the default chain count; nsim not divisible by n.chains
    assert lqg.tmvrnorm(obj, 7, {"burn.in": 2, "seed": 4}, method="HMC").shape == (3, 7)
    assert lqg.tmvrnorm(obj, 10, {"burn.in": 2, "n.chains": 3, "seed": 4}, method="HMC").shape == (3, 10)
    # the reference's layout: ONE serial chain (n.chains = 1), burn.in = 0
    x1 = lqg.tmvrnorm(obj, 300, {"burn.in": 0, "n.chains": 1, "seed": 5}, method="HMC")
    assert lqg.last_stats["transitions"] == 300 and x1[1].min() >= 0.1


def test_canonical_without_constraints_and_restart_cap(lqg):
    """numlin = 0: every transition returns the drawn momentum (T = pi/2).  An impossible
    geometry trips max_restarts and is reported as a status code instead of tmg's endless loop."""
    from lineqgpr_b200 import _lib as L
    d = 5
    draws = np.random.default_rng(0).standard_normal(4 * d)
    z, st = lqg.rtmg_canonical(4, 0, np.zeros(d), np.zeros((0, d)), np.zeros(0), injected=draws)
    assert st["bounces"] == 0
    assert np.allclose(z.T, draws.reshape(4, d), rtol=0, atol=1e-15)
    # infeasible start whose orbit (amplitude ~40) never reaches the region z_0 >= 50: every end
    # point fails _verifyConstraints, tmg would loop forever (tmg:src/HmcSampler.cpp:51,130)
    F = np.array([[1.0, 0.0]]); g = np.array([-50.0])
    with pytest.raises(L.LqgError) as e:
        lqg.rtmg_canonical(2, 1, np.array([40.0, 0.0]), F, g, max_restarts=20)
    assert e.value.status == 5
    # the far tail itself (start inside z_0 >= 50) is sampled fine
    z, st = lqg.rtmg_canonical(50, 1, np.array([50.5, 0.0]), F, g)
    assert z[0].min() >= 50.0 and st["bounces"] > 0


@pytest.mark.gpu
def test_guard_band_mode_finds_no_out_of_bounds_writes():
    """LQG_GUARD=1 puts canary bands around every internal device buffer; a pass over every kernel
    family at ragged small sizes must leave all of them intact (the library's own bounds check)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, LQG_GUARD="1")
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "sanitize_small.py")], cwd=root, env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "guard mode 1 violations 0" in r.stdout
    assert "[lqg guard]" not in r.stderr



def test_canonical_lockstep_rounds_follow_the_per_chain_kernel(lqg):
    """lqg_opts.canon_mode = 2 (one tensor-core GEMM per hit-search round for all chains) against
    mode 1 (one persistent CTA per chain): same Philox draws -> the same chains up to the rounding
    of the dot products, same event counts on short runs; many chains with different lengths
    (burn-in, ragged finish), c < d and c > d, restart cap, injected-draw exhaustion."""
    from lineqgpr_b200 import _lib
    rng = np.random.default_rng(5)
    for d, c, K, n in ((30, 12, 7, 5), (24, 90, 70, 4), (200, 450, 96, 3)):
        F = rng.standard_normal((c, d)) / np.sqrt(d)
        g = np.full(c, 1.5)
        z0 = np.zeros(d)
        z1, s1 = lqg.rtmg_canonical(n, 9, z0, F, g, n_chains=K, burn_in=2, mode=1)
        z2, s2 = lqg.rtmg_canonical(n, 9, z0, F, g, n_chains=K, burn_in=2, mode=2)
        assert s2["violations"] == 0 and s2["transitions"] == K * (n + 2) == s1["transitions"]
        assert (F @ z2 + g[:, None]).min() >= -1e-12
        assert abs(s1["bounces"] - s2["bounces"]) <= max(2, s1["bounces"] // 200), (s1["bounces"], s2["bounces"])
        close = np.abs(z1 - z2).max(axis=0) <= 1e-8 * max(1.0, np.abs(z1).max())
        assert close.mean() >= 0.97, close.mean()          # a rounding-level tie in a hit search may fork a chain
        assert s2["hit_searches"] >= s2["bounces"] + s2["transitions"]
    # errors are reported the same way
    F = np.array([[1.0, 0.0], [-1.0, 0.0]]); g = np.array([-50.0, 60.0])        # start (40, 0) violates z0 >= 50
    with pytest.raises(_lib.LqgError, match="RESTART_CAP"):
        lqg.rtmg_canonical(2, 1, np.array([40.0, 0.0]), F, g, n_chains=3, max_restarts=20, mode=2)
    with pytest.raises(_lib.LqgError, match="DRAWS_EXHAUSTED"):
        lqg.rtmg_canonical(5, 1, np.zeros(4), np.eye(4), np.ones(4), n_chains=2, injected=np.zeros((2, 8)), mode=2)


def test_packed_and_per_warp_exact_evaluation_are_bitwise_identical(lqg):
    """The packed evaluation of the candidates of all warps (default from 4 warps per chain) and
    the per-warp evaluation (LQG_BOX_VARIANT=n) select the same walls at the same times: the
    arg-min with the first-index tie break does not depend on who evaluated what.  Compared in
    two fresh processes (the variant is read once per process) at d = 300, 600 and 1100."""
    import hashlib
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys, hashlib, numpy as np; sys.path.insert(0, '.');"
            "import lineqgpr_b200 as lqg; from lineqgpr_b200 import workloads as W;"
            "h = hashlib.sha256();"
            "[h.update(np.ascontiguousarray(lqg.tmvrnorm(dict(W.cfg4(m=m)[0], **{'class': 'HMC'}), 96, "
            "{'burn.in': 3, 'n.chains': 12, 'seed': 4, 'whiten': 'ul'})).tobytes()) for m in (150, 300, 550)];"
            "print('HASH', h.hexdigest())")
    out = {}
    for variant in ("", "n"):
        env = dict(os.environ)
        env.pop("LQG_BOX_VARIANT", None)
        if variant:
            env["LQG_BOX_VARIANT"] = variant
        r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        out[variant] = [l for l in r.stdout.splitlines() if l.startswith("HASH")][0]
    assert out[""] == out["n"]
