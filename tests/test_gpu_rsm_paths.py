"""GPU parity tests of RSM, the FP64 tensor-core GEMM and the band summaries (through the C ABI)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_rsm_decisions_match_oracle_golden(lqg, golden):
    """Same injected normals and uniforms -> identical accept/reject decisions, same columns."""
    G = golden("rsm_cfg2")
    nsim = G["xi"].shape[1]
    obj = dict(mu=G["map"], map=G["map"], Sigma=G["factor"] @ G["factor"].T, lb=G["lb"], ub=G["ub"])
    x = lqg.tmvrnorm(obj, nsim, {"factor": G["factor"], "injected": G["Z"], "injected.u": G["U"]}, method="RSM")
    st = dict(lqg.last_stats)
    assert np.abs(x - G["xi"]).max() < 1e-12
    assert st["accepted"] == nsim
    assert st["proposals"] == int(G["accepted_idx"][-1]) + 1
    assert st["in_box"] == int(G["inbox"][:st["proposals"]].sum())


def test_rsm_fewer_samples_is_a_prefix(lqg, golden):
    G = golden("rsm_cfg2")
    obj = dict(mu=G["map"], map=G["map"], Sigma=G["factor"] @ G["factor"].T, lb=G["lb"], ub=G["ub"])
    x = lqg.tmvrnorm(obj, 10, {"factor": G["factor"], "injected": G["Z"], "injected.u": G["U"]}, method="RSM")
    assert np.abs(x - G["xi"][:, :10]).max() < 1e-12
    assert lqg.last_stats["proposals"] == int(G["accepted_idx"][9]) + 1


def test_rsm_exhausted_draws_is_an_error(lqg, golden):
    from lineqgpr_b200 import _lib as L
    G = golden("rsm_cfg2")
    obj = dict(mu=G["map"], map=G["map"], Sigma=G["factor"] @ G["factor"].T, lb=G["lb"], ub=G["ub"])
    with pytest.raises(L.LqgError) as e:
        lqg.tmvrnorm(obj, 500, {"factor": G["factor"], "injected": G["Z"], "injected.u": G["U"]}, method="RSM")
    assert e.value.status == 6


def test_rsm_philox_statistics(lqg, oracle):
    """Philox mode on cfg2: all samples in the box; moments agree with an oracle run."""
    from lineqgpr_b200 import workloads as W
    par, _ = W.cfg2(sampler="RSM")
    n = 4000
    x = lqg.tmvrnorm(par, n, {"seed": 5})
    st = dict(lqg.last_stats)
    assert x.shape == (20, n) and st["accepted"] == n and st["proposals"] > st["in_box"] >= n
    assert np.all(x >= par["lb"][:, None]) and np.all(x <= par["ub"][:, None])
    rng = np.random.default_rng(0)
    P = int(st["proposals"] * 1.3)
    xo, info = oracle.tmvrnorm_RSM(par, n, rng.standard_normal((20, P)), rng.uniform(size=P))
    assert xo.shape[1] == n
    se = np.sqrt((x.var(1) + xo.var(1)) / n)
    assert (np.abs(x.mean(1) - xo.mean(1)) > 3.5 * se).sum() <= 1
    rate_g = n / st["proposals"]; rate_o = n / info["n_proposals"]
    assert abs(rate_g - rate_o) < 5 * np.sqrt(rate_o / info["n_proposals"]) + 0.002


@pytest.mark.parametrize("shape", [(1, 1, 1), (7, 5, 3), (128, 64, 16), (130, 67, 19), (300, 1000, 300), (1000, 257, 1000)])
def test_dgemm_matches_numpy(lqg, shape):
    m, n, k = shape
    rng = np.random.default_rng(m * 31 + n)
    A = rng.standard_normal((m, k)); B = rng.standard_normal((k, n))
    C, rs = lqg.dgemm(A, B, row_sum=True)
    ref = A @ B
    tol = 1e-13 * k * max(1.0, np.abs(A).max() * np.abs(B).max())
    assert np.abs(C - ref).max() < tol
    assert np.abs(rs - ref.sum(1)).max() < tol * n


@pytest.mark.parametrize("shape", [(130, 67, 20, 130, 20), (129, 70, 40, 132, 44), (256, 130, 48, 256, 48),
                                   (1000, 258, 1000, 1000, 1000), (333, 64, 16, 400, 18), (96, 8, 64, 96, 65)])
def test_dgemm_leading_dimensions_and_ring_edges(lqg, shape):
    """lqg_dgemm on sub-matrices (lda > m, ldb > k): even leading dimensions take the bulk-copy ring kernel
    (full tiles by TMA, partial k-tiles / odd row counts staged by the threads), odd ones the
    register-staged kernel; both must agree with numpy."""
    from lineqgpr_b200 import _lib as L
    m, n, k, lda, ldb = shape
    rng = np.random.default_rng(m + 7 * n + k)
    Abig = np.asfortranarray(rng.standard_normal((lda, k))); Bbig = np.asfortranarray(rng.standard_normal((ldb, n)))
    Abig[m:, :] = np.nan; Bbig[k:, :] = np.nan            # rows outside the operands must never be used
    ldc = m + 3
    C = np.full((ldc, n), -7.0, order="F"); rs = np.zeros(m)
    rc = L.lib().lqg_dgemm(m, n, k, L.ptr(Abig), lda, L.ptr(Bbig), ldb, L.ptr(C), ldc, L.ptr(rs),
                           L.make_opts(mem=L.MEM_HOST), None)
    L.check(rc, "lqg_dgemm")
    ref = Abig[:m] @ Bbig[:k]
    tol = 1e-13 * k * max(1.0, np.abs(ref).max())
    assert np.abs(C[:m] - ref).max() < tol
    assert np.abs(rs - ref.sum(1)).max() < tol * n


def test_simulate_tail_matches_oracle(lqg, oracle):
    """xi.sim = qr.solve(Lambda, eta); ysim = Phi.test %*% xi.sim  (R/lineqGP.R:639-640)."""
    from lineqgpr_b200 import workloads as W
    from lineqgpr_b200.simulate import lambda_pinv
    par, aux = W.cfg4(m=40)                  # tall Lambda (80 x 40): least-squares branch (Q8)
    rng = np.random.default_rng(1)
    eta = rng.standard_normal((80, 333))
    xi = lqg.dgemm(lambda_pinv(aux["Lambda"]), eta)
    ysim = lqg.dgemm(aux["Phi_test"], xi)
    xi_o, y_o = oracle.simulate_tail(aux["Lambda"], aux["Phi_test"], eta)
    assert np.abs(xi - xi_o).max() < 1e-10 * max(1, np.abs(xi_o).max())
    assert np.abs(ysim - y_o).max() < 1e-10 * max(1, np.abs(y_o).max())


def test_bands_golden_and_random(lqg, oracle, golden):
    G = golden("bands")
    mean, q = lqg.bands(G["y"], G["probs"])
    assert np.allclose(mean, G["mean"], rtol=1e-14, atol=1e-15)
    assert np.array_equal(q, G["q"])                      # order statistics are exact
    rng = np.random.default_rng(9)
    for (n_t, N) in [(1, 1), (3, 2), (5, 10), (33, 4097), (200, 10000)]:
        y = rng.standard_normal((n_t, N)) * 10.0 ** rng.integers(-3, 4, size=(n_t, 1))
        y[0, : N // 2] = y[0, 0]                           # ties
        probs = np.array([0.0, 0.025, 0.05, 0.5, 0.95, 0.975, 1.0])
        mean, q = lqg.bands(y, probs)
        mo, qo = oracle.bands(y, probs)
        assert np.allclose(mean, mo, rtol=1e-13, atol=1e-300)
        assert np.array_equal(q, qo), (n_t, N)


def test_simulate_end_to_end_cfg1(lqg):
    """create -> simulate (HMC) -> bands on config 1: constraints hold for eta and for xi
    (square Lambda), sample paths are monotone at the knots."""
    from lineqgpr_b200 import workloads as W
    par, aux = W.cfg1()
    sim = lqg.simulate(aux["lineq_model"], nsim=1000, seed=1, xtest=aux["xtest"])
    xi = sim["xi.sim"]
    assert xi.shape == (100, 1000) and sim["ysim"].shape == (100, 1000)
    assert np.diff(xi, axis=0).min() > -1e-9
    assert sim["stats"]["violations"] == 0
    mean, q = lqg.bands(sim["ysim"])
    assert np.all(q[0] <= mean + 1e-12) and np.all(mean <= q[1] + 1e-12)
    # posterior mean path follows the sigmoid data
    assert abs(mean[0]) < 0.2 and abs(mean[-1] - 1.0) < 0.2


def test_paths_bands_slabs_match_materialised(lqg, oracle):
    """Slab-wise Phi xi -> bands on device memory == bands of the materialised ysim (cfg5-like, small)."""
    from lineqgpr_b200 import workloads as W
    par, aux = W.cfg5(D=8, knots=10, n=80, n_test=333)
    rng = np.random.default_rng(2)
    xi = rng.standard_normal((80, 2501))
    Phi = aux["Phi_test"]
    probs = (0.025, 0.5, 0.975)
    mean, q, info = lqg.paths_bands(Phi, xi, probs, slab_rows=100)
    y = Phi @ xi
    mo, qo = oracle.bands(y, probs)
    assert np.allclose(mean, mo, rtol=1e-11, atol=1e-13)
    assert np.allclose(q, qo, rtol=1e-11, atol=1e-13)       # GEMM rounding differs from numpy's, selection is exact
    assert info["gemm_ms"] > 0 and info["bands_ms"] > 0


def test_bands_bracketed_path_is_exact(lqg, oracle):
    """The one-pass bracketed path (sampled brackets -> collect -> select) against the oracle's
    sort-based type-7 quantiles, bit for bit: ragged tile counts, heavy ties (bracket overflow ->
    per-row exact fallback), an adversarial row whose systematic sample is unrepresentative
    (bracket miss -> fallback), constant slabs (every row falls back -> whole-slab exact path),
    extreme probabilities."""
    rng = np.random.default_rng(21)
    probs = np.array([0.0, 0.001, 0.025, 0.5, 0.75, 0.975, 0.9999, 1.0])
    for (n_t, N) in [(70, 50000), (33, 8192), (1, 100001), (97, 20011)]:
        y = rng.standard_normal((n_t, N)) * 10.0 ** rng.integers(-2, 3, size=(n_t, 1)) + rng.standard_normal((n_t, 1))
        y[0, : N // 2] = y[0, 0]                                   # half the row tied
        if n_t > 2:
            y[1] = np.round(y[1])                                  # few distinct values: massive ties
            stride = N / 2048.0                                    # the sampled positions see only huge values
            idx = np.unique((np.arange(2048) * stride).astype(np.int64))
            y[2, idx] = 1e6 + rng.standard_normal(idx.size)
        mean, q = lqg.bands(y, probs)
        mo, qo = oracle.bands(y, probs)
        assert np.array_equal(q, qo), (n_t, N)
        assert np.allclose(mean, mo, rtol=1e-12, atol=1e-300)
    y = np.full((100, 9000), 3.25)                                 # every bracket overflows
    mean, q = lqg.bands(y, probs[:3])
    assert np.array_equal(q, np.full((3, 100), 3.25)) and np.array_equal(mean, np.full(100, 3.25))
    mean, q = lqg.bands(rng.standard_normal((40, 30000)), np.zeros(0))   # means only
    assert q.shape[0] == 0 and mean.shape == (40,)


def test_lqg_paths_one_call_matches_gemm_plus_bands(lqg, oracle):
    """lqg_paths (native slab loop: GEMM -> exact bands -> optional ysim copy-out) against numpy + the
    oracle's type-7 quantiles: several slabs with a ragged last one, with and without ysim."""
    rng = np.random.default_rng(12)
    Phi = rng.uniform(size=(301, 40)); xi = rng.standard_normal((40, 9000))
    probs = np.array([0.025, 0.5, 0.975])
    y = Phi @ xi
    mo, qo = oracle.bands(y, probs)
    for slab in (0, 128, 77):
        mean, q, info = lqg.paths(Phi, xi, probs, slab_rows=slab)
        assert np.allclose(mean, mo, rtol=1e-12, atol=1e-13)
        assert np.allclose(q, qo, rtol=1e-12, atol=1e-12)        # GEMM summation order differs from numpy's by rounding
        assert info["gemm_ms"] > 0 and info["bands_ms"] > 0
    mean, q, ysim, _ = lqg.paths(Phi, xi, probs, return_ysim=True, slab_rows=100)
    assert np.allclose(ysim, y, rtol=1e-12, atol=1e-12)
    m2, q2 = lqg.bands(ysim, probs)
    assert np.array_equal(q, q2) and np.allclose(mean, m2, rtol=1e-14)
    mean0, q0, _ = lqg.paths(Phi, xi, np.zeros(0))
    assert q0.shape == (0, 301) and np.allclose(mean0, mo, rtol=1e-12, atol=1e-13)


def test_lqg_paths_identity_slabs_take_every_fallback_route(lqg, oracle):
    """Several slabs of a sparse Phi (ELL product, bands of the stored slab): slabs whose test points go through the
    per-row exact path, a slab that goes through the whole-slab exact path and clean slabs, interleaved, against the
    oracle bit for bit (Phi = I, so ysim is the crafted matrix itself)."""
    rng = np.random.default_rng(5)
    n_t, N = 400, 9001
    y = rng.standard_normal((n_t, N)) * 10.0 ** rng.integers(-2, 3, size=(n_t, 1))
    y[85:160] = 3.25                                               # slab 1 (rows 80..159): 75 constant rows -> whole slab
    y[170, : N // 2] = y[170, 0]                                   # slab 2: two rows with heavy ties -> single-row launches
    y[200] = np.round(y[200])
    y[399] = np.round(y[399])                                      # last slab
    probs = np.array([0.025, 0.5, 0.975])
    mo, qo = oracle.bands(y, probs)
    mean, q, info = lqg.paths(np.eye(n_t), y, probs, slab_rows=80)
    assert np.array_equal(q, qo)
    assert np.allclose(mean, mo, rtol=1e-12, atol=1e-300)
    mean1, q1, _ = lqg.paths(np.eye(n_t), y, probs, slab_rows=n_t)
    assert np.array_equal(q, q1) and np.allclose(mean, mean1, rtol=1e-14, atol=1e-300)   # split count follows the slab size


def _dense_case(seed, n_t=300, m=40, N=20011):
    rng = np.random.default_rng(seed)
    return rng.uniform(size=(n_t, m)), rng.standard_normal((m, N)), np.array([0.025, 0.5, 0.975])


@pytest.mark.parametrize("slab", [0, 128, 76, 75])
def test_fused_product_bands_equal_bands_of_the_stored_product(lqg, oracle, slab):
    """Dense Phi, >= 8192 samples, paths not requested: the product is summarised in its own epilogue and never
    stored.  The order statistics must be those of the very same product (same kernel, stored this time) bit for
    bit, and agree with numpy's product within GEMM rounding; slabs of 75 rows start on odd rows every other
    time, which the tensor-map kernel cannot take: those go through the stored form."""
    Phi, xi, probs = _dense_case(31)
    mean, q, info = lqg.paths(Phi, xi, probs, slab_rows=slab)
    mean_s, q_s, ysim, _ = lqg.paths(Phi, xi, probs, return_ysim=True, slab_rows=slab)
    mo, qo = oracle.bands(ysim, probs)
    assert np.array_equal(q_s, qo)
    assert np.array_equal(q, qo)
    tol = 1e-14 * np.abs(ysim).max()                                       # rows cancel: compare on the scale of the values summed
    assert np.abs(mean - mo).max() <= tol and np.abs(mean_s - mo).max() <= tol
    m2, q2 = oracle.bands(Phi @ xi, probs)
    assert np.allclose(q, q2, rtol=1e-12, atol=1e-12) and np.allclose(mean, m2, rtol=1e-12, atol=1e-13)
    assert info["gemm_ms"] > 0 and info["bands_ms"] > 0
    pe = np.array([0.0, 0.0001, 0.2, 0.9999, 1.0])                          # open-ended brackets
    _, qe, _ = lqg.paths(Phi, xi, pe, slab_rows=slab)
    assert np.array_equal(qe, oracle.bands(ysim, pe)[1])
    mean0, q0, _ = lqg.paths(Phi, xi, np.zeros(0), slab_rows=slab)          # means only
    assert q0.shape == (0, Phi.shape[0]) and np.abs(mean0 - mo).max() <= tol


def test_fused_product_bands_fallback_routes(lqg, oracle):
    """Test points the bracketed statistics cannot settle inside the fused product: a few rows with massive ties
    (their row of the product is formed again by the same kernel, odd and even rows, first and last of a slab)
    and slabs where every row is tied (the slab is stored after all) — bit for bit against the oracle on the
    stored product."""
    Phi, xi, probs = _dense_case(32)
    n_t = Phi.shape[0]
    xi[7] = np.round(xi[7])                                        # rows of Phi that pick coordinate 7 see few distinct values
    for r in (0, 3, 10, 149, 150, 224, 298, 299):
        Phi[r] = 0.0; Phi[r, 7] = 1.0
    for slab in (0, 75, 150):
        mean, q, _ = lqg.paths(Phi, xi, probs, slab_rows=slab)
        mean_s, q_s, ysim, _ = lqg.paths(Phi, xi, probs, return_ysim=True, slab_rows=slab)
        mo, qo = oracle.bands(ysim, probs)
        assert np.array_equal(q, qo), slab
        assert np.abs(mean - mo).max() <= 1e-14 * np.abs(ysim).max()
    xi2 = xi.copy()
    xi2[:, : xi.shape[1] // 2] = xi2[:, [0]]                       # half of the samples identical: every row is half tied
    mean, q, _ = lqg.paths(Phi, xi2, probs, slab_rows=150)
    _, _, ysim, _ = lqg.paths(Phi, xi2, probs, return_ysim=True, slab_rows=150)
    _, qo = oracle.bands(ysim, probs)
    mo = np.asarray(ysim, np.longdouble).mean(axis=1).astype(float)   # 10 005 equal terms: a plain running sum drifts by 1e-12
    assert np.array_equal(q, qo)
    assert np.abs(mean - mo).max() <= 1e-14 * np.abs(ysim).max()
