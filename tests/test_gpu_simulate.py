"""GPU parity tests of lqg_simulate: the whole simulate tail in one C-ABI call must equal the same
tail run stage by stage (lqg_hmc_box -> lqg_dgemm -> lqg_dgemm -> lqg_bands) and the oracle's
restatement of R/lineqGP.R:611-640 / R/lineqGPutils.R:422-423."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def _cfg(d_case):
    from lineqgpr_b200 import workloads as W
    if d_case == "cfg1":                                   # d = m = 100: one launch, no pool rounds
        par, aux = W.cfg1()
    else:
        par, aux = W.cfg4(m=150, n_test=131)               # d = 300 = 2 m (tall Lambda), pool rounds
    return par, aux["Lambda"], aux["Phi_test"]


@pytest.mark.parametrize("case", ["cfg1", "cfg4small"])
def test_pipeline_equals_stage_by_stage(lqg, oracle, case):
    from lineqgpr_b200.simulate import simulate_pipeline, lambda_pinv, dgemm, bands
    par, Lam, Phi = _cfg(case)
    ctl = {"burn.in": 5, "n.chains": 24, "seed": 17, "burn.in.exact": True}
    nsim = 24 * 40
    probs = (0.025, 0.5, 0.975)
    r = simulate_pipeline(par, nsim, Lambda=Lam, Phi_test=Phi, probs=probs, control=ctl, return_eta=True, return_ysim=True)
    assert r["stats"]["violations"] == 0 and r["stats"]["n_total"] == nsim
    # (1) the sampler is lqg_hmc_box: same chains, same Philox streams
    eta = lqg.tmvrnorm(dict(par, **{"class": "HMC"}), nsim, ctl)
    assert np.array_equal(r["eta"], eta)
    # (2) xi = qr.solve(Lambda, eta)
    xi_ref = np.linalg.lstsq(Lam, eta, rcond=None)[0]                      # qr.solve(Lambda, eta), R/lineqGP.R:639
    assert np.abs(r["xi.sim"] - xi_ref).max() <= 1e-10 * max(1.0, np.abs(xi_ref).max())
    # (3) ysim = Phi xi, columns in the order of xi.sim
    y_ref = Phi @ r["xi.sim"]
    assert r["ysim"].shape == y_ref.shape
    assert np.abs(r["ysim"] - y_ref).max() <= 1e-12 * max(1.0, np.abs(y_ref).max())
    # (4) bands: exact order statistics of those very paths, and the oracle's type-7 rule
    mo, qo = oracle.bands(r["ysim"], probs)
    assert np.array_equal(r["qtls"], qo)
    assert np.abs(r["ymean"] - mo).max() <= 1e-14 * max(1.0, np.abs(mo).max())
    mean_g, q_g = bands(r["ysim"], probs)
    assert np.array_equal(q_g, r["qtls"])


def test_pipeline_device_basis_equals_uploaded_phi(lqg):
    """Phi built on the device from (xtest, knots) = the uploaded hat-basis matrix, bit for bit."""
    from lineqgpr_b200 import workloads as W
    from lineqgpr_b200.simulate import simulate_pipeline
    par, aux = W.cfg1()
    ctl = {"burn.in": 3, "n.chains": 16, "seed": 5, "burn.in.exact": True}
    u = np.asarray(aux["model"]["localParam"]["u"], float).ravel()
    a = simulate_pipeline(par, 320, Lambda=aux["Lambda"], Phi_test=aux["Phi_test"], control=ctl)
    b = simulate_pipeline(par, 320, Lambda=aux["Lambda"], xtest=np.asarray(aux["xtest"], float), ulist=[u], control=ctl)
    assert np.array_equal(a["xi.sim"], b["xi.sim"])
    assert np.array_equal(a["qtls"], b["qtls"]) and np.array_equal(a["ymean"], b["ymean"])


def test_pipeline_identity_lambda_and_no_paths(lqg):
    """d == m without Lambda: xi = eta; n_t = 0 stops after xi."""
    from lineqgpr_b200.simulate import simulate_pipeline
    rng = np.random.default_rng(3)
    d = 40
    B = rng.standard_normal((d, d)); S = B @ B.T / d + 0.3 * np.eye(d)
    par = dict(mu=np.zeros(d), Sigma=S, lb=np.full(d, -0.5), ub=np.full(d, 1.5))
    r = simulate_pipeline(par, 200, control={"burn.in": 10, "n.chains": 8, "seed": 2, "burn.in.exact": True}, return_eta=True)
    assert np.array_equal(r["xi.sim"], r["eta"]) and r["ymean"] is None
    assert np.all(r["eta"] >= -0.5) and np.all(r["eta"] <= 1.5)


def test_pipeline_slabs_and_ragged_rows(lqg, oracle):
    """Several row slabs (slab_rows smaller than n_t, not a divisor) give the same bands."""
    from lineqgpr_b200.simulate import simulate_pipeline
    par, Lam, Phi = _cfg("cfg1")
    ctl = {"burn.in": 3, "n.chains": 8, "seed": 9, "burn.in.exact": True}
    a = simulate_pipeline(par, 400, Lambda=Lam, Phi_test=Phi, control=ctl)
    b = simulate_pipeline(par, 400, Lambda=Lam, Phi_test=Phi, control=ctl, slab_rows=37)
    assert np.array_equal(a["qtls"], b["qtls"])
    assert np.abs(a["ymean"] - b["ymean"]).max() <= 1e-15 * max(1.0, np.abs(a["ymean"]).max())


def test_pipeline_fused_bands_equal_bands_of_stored_paths(lqg, oracle):
    """Dense Phi and enough samples for the pilot product to pay (N > 2048 m / 22): without ysim_out the product is
    summarised in its epilogue and never stored; with it the slab is stored and read again.  Same order statistics
    bit for bit, also slab by slab."""
    from lineqgpr_b200.simulate import simulate_pipeline
    par, Lam, _ = _cfg("cfg4small")
    rng = np.random.default_rng(8)
    Phi = rng.uniform(size=(130, Lam.shape[1]))
    ctl = {"burn.in": 5, "n.chains": 24, "seed": 3, "burn.in.exact": True}
    probs = (0.025, 0.5, 0.975)
    a = simulate_pipeline(par, 24 * 600, Lambda=Lam, Phi_test=Phi, probs=probs, control=ctl)
    b = simulate_pipeline(par, 24 * 600, Lambda=Lam, Phi_test=Phi, probs=probs, control=ctl, return_ysim=True)
    c = simulate_pipeline(par, 24 * 600, Lambda=Lam, Phi_test=Phi, probs=probs, control=ctl, slab_rows=48)
    assert np.array_equal(a["xi.sim"], b["xi.sim"])
    mo, qo = oracle.bands(b["ysim"], probs)
    assert np.array_equal(b["qtls"], qo)
    assert np.array_equal(a["qtls"], qo) and np.array_equal(c["qtls"], qo)
    tol = 1e-14 * np.abs(b["ysim"]).max()
    assert np.abs(a["ymean"] - mo).max() <= tol and np.abs(c["ymean"] - mo).max() <= tol


@pytest.mark.parametrize("lam_kind,phi_kind", [("banded2", "hat"), ("banded2", "dense"), ("dense", "hat"), ("blockdiag", "hat3")])
def test_pipeline_structured_and_dense_products(lqg, lam_kind, phi_kind):
    """xi = qr.solve(Lambda, eta) and ysim = Phi xi take structure-aware kernels when Lambda's rows have short
    non-zero windows (banded normal equations + refinement) and Phi's rows few non-zeros (ELL product), and the
    dense tensor-core products otherwise: every combination must agree with numpy's least squares / product."""
    from lineqgpr_b200.simulate import simulate_pipeline
    rng = np.random.default_rng(11)
    m = 70
    I = np.eye(m); D1 = np.diff(I, axis=0); D2 = np.diff(I, n=2, axis=0)
    if lam_kind == "banded2":
        Lam = np.vstack([I, 3.0 * D1, -0.5 * D2])            # bounded + monotone + convex rows: windows of 1, 2, 3 columns
    elif lam_kind == "blockdiag":
        blk = np.vstack([np.eye(10), np.diff(np.eye(10), axis=0)])
        Lam = np.kron(np.eye(7), blk)                          # additive model: one block per input dimension
    else:
        Lam = rng.standard_normal((2 * m, m))                  # no structure: dense Lambda^+ product
    d = Lam.shape[0]
    B = rng.standard_normal((d, d)); S = B @ B.T / d + 0.5 * np.eye(d)
    par = dict(mu=0.1 * rng.standard_normal(d), Sigma=S, lb=np.full(d, -2.0), ub=np.full(d, 2.5))
    n_t = 53
    if phi_kind == "dense":
        Phi = rng.standard_normal((n_t, m))
    else:
        Phi = np.zeros((n_t, m))
        for t in range(n_t):                                    # hat functions: 2 neighbours (hat3: a third entry further away)
            j = rng.integers(0, m - 1); w = rng.uniform()
            Phi[t, j], Phi[t, j + 1] = 1 - w, w
            if phi_kind == "hat3":
                Phi[t, (j + 17) % m] += 0.25
    ctl = {"burn.in": 4, "n.chains": 12, "seed": 23, "burn.in.exact": True}
    r = simulate_pipeline(par, 12 * 50, Lambda=Lam, Phi_test=Phi, control=ctl, return_eta=True, return_ysim=True)
    xi_ref = np.linalg.lstsq(Lam, r["eta"], rcond=None)[0]
    assert np.abs(r["xi.sim"] - xi_ref).max() <= 1e-11 * max(1.0, np.abs(xi_ref).max())
    y_ref = Phi @ r["xi.sim"]
    assert np.abs(r["ysim"] - y_ref).max() <= 1e-12 * max(1.0, np.abs(y_ref).max())
    assert np.abs(r["ymean"] - r["ysim"].mean(1)).max() <= 1e-13 * max(1.0, np.abs(y_ref).max())


def test_simulate_lineqGP_uses_the_pipeline(lqg):
    """simulate() (R/lineqGP.R:601-643) routes HMC models through lqg_simulate: one call, and its
    result is what the stage-by-stage path returns."""
    from lineqgpr_b200 import workloads as W
    par, aux = W.cfg1()
    ctl = {"n.chains": 16, "seed": 4}
    a = lqg.simulate(aux["lineq_model"], nsim=160, seed=1, xtest=aux["xtest"], control=ctl)
    b = lqg.simulate(aux["lineq_model"], nsim=160, seed=1, xtest=aux["xtest"], control=dict(ctl, pipeline="stages"))
    assert "ymean" in a and "ymean" not in b
    assert np.array_equal(a["eta"], b["eta"])
    assert np.abs(a["xi.sim"] - b["xi.sim"]).max() <= 1e-11 * np.abs(b["xi.sim"]).max()
    assert np.abs(a["ysim"] - b["ysim"]).max() <= 1e-11 * np.abs(b["ysim"]).max()


def test_pipeline_errors(lqg):
    from lineqgpr_b200 import _lib as L
    from lineqgpr_b200.simulate import simulate_pipeline
    par = dict(mu=np.zeros(3), Sigma=np.eye(3), lb=-np.ones(3), ub=np.ones(3))
    with pytest.raises(L.LqgError) as e:      # d != m and no Lambda
        a = L.SimulateArgs(); a.d, a.m = 3, 2
        v = np.zeros(3); S = L.f64(np.eye(3)); lo = -np.ones(3); hi = np.ones(3)
        a.mu, a.Sigma, a.lb, a.ub, a.init = L.ptr(v), L.ptr(S), L.ptr(lo), L.ptr(hi), L.ptr(v)
        a.n_per_chain = 2
        L.check(L.lib().lqg_simulate(L.C.byref(a), L.make_opts(), None, None), "lqg_simulate")
    assert e.value.status == 1
    with pytest.raises(ValueError):
        simulate_pipeline(dict(par, lb=np.ones(3)), 4)
    bad = dict(par, Sigma=np.array([[1.0, 2, 0], [2, 1, 0], [0, 0, 1]]))
    with pytest.raises(ValueError, match="positive definite"):
        simulate_pipeline(bad, 4, control={"n.chains": 2})


def test_two_ranks_equal_one_process():
    """Sharded job (2 ranks, NCCL all-gather of the xi shards and of the band pieces) == the same job in
    one process, bit for bit (scripts/dist_simulate_check.py).  Needs two GPUs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", str(ROOT / "scripts" / "dist_simulate_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "DIST_SIMULATE_OK" in r.stdout
