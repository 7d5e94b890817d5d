"""simulate.lineqGP / simulate.lineqAGP tails and the band summaries, with the per-sample work
on the GPU.  Mirrors

  simulate.lineqGP    R/lineqGP.R:601-643   (eta-space parameters :611-614, tmvrnorm :626,
                                             xi.sim = qr.solve(Lambda, eta) :639, ysim :640)
  simulate.lineqAGP   R/lineqAGP.R:520-596  (mu := eta_MAP :567 (Q4), tmvrnorm :573, qr.solve :575;
                                             PhiAll.test %*% xiAll.sim is left to the user :509-512)
  band summaries      R/lineqGPutils.R:422-423, :294-296
"""
from __future__ import annotations

import time

import numpy as np

from . import _lib as L
from . import model as M
from .samplers import tmvrnorm, last_stats


def dgemm(A, B, row_sum=False, want_C=True):
    """C = A %*% B on the FP64 tensor cores (lqg_dgemm). Host arrays in, host arrays out."""
    A = L.f64(np.atleast_2d(np.asarray(A, float))); B = L.f64(np.atleast_2d(np.asarray(B, float)))
    m, k = A.shape
    k2, n = B.shape
    if k != k2:
        raise ValueError("non-conformable arguments")
    C = np.empty((m, n), order="F") if want_C else None
    rs = np.zeros(m) if row_sum else None
    ms = L.C.c_double(0.0)
    rc = L.lib().lqg_dgemm(m, n, k, L.ptr(A), m, L.ptr(B), k, L.ptr(C), m, L.ptr(rs),
                           L.make_opts(mem=L.MEM_HOST), L.C.byref(ms))
    L.check(rc, "lqg_dgemm")
    return (C, rs) if row_sum else C


def bands(ysim, probs=(0.025, 0.975)):
    """(ymean, qtls) = (apply(ysim,1,mean), apply(ysim,1,quantile,probs)); qtls is nprobs x n_t."""
    ysim = L.f64(np.atleast_2d(np.asarray(ysim, float)))
    n_t, N = ysim.shape
    probs = np.ascontiguousarray(probs, dtype=np.float64)
    mean = np.empty(n_t)
    q = np.empty((probs.size, n_t), order="F")
    ms = L.C.c_double(0.0)
    rc = L.lib().lqg_bands(n_t, N, L.ptr(ysim), n_t, L.ptr(probs), probs.size, L.ptr(mean), L.ptr(q),
                           L.make_opts(mem=L.MEM_HOST), L.C.byref(ms))
    L.check(rc, "lqg_bands")
    return mean, q


def paths(Phi_test, xi_sim, probs=(0.025, 0.975), return_ysim=False, slab_rows=0):
    """ysim = Phi.test %*% xi.sim with its mean and quantile bands through ONE C-ABI call
    (lqg_paths): the slab loop (tensor-core GEMM -> exact bands -> optional copy-out) runs inside
    the library, host arrays in and out, no torch.  Returns (ymean, qtls[, ysim], info)."""
    Phi = L.f64(np.atleast_2d(np.asarray(Phi_test, float)))
    xi = L.f64(np.atleast_2d(np.asarray(xi_sim, float)))
    n_t, m = Phi.shape
    if xi.shape[0] != m:
        raise ValueError("non-conformable arguments")
    N = xi.shape[1]
    probs = np.ascontiguousarray(probs, dtype=np.float64)
    mean = np.empty(n_t)
    q = np.empty((probs.size, n_t), order="F")
    ysim = np.empty((n_t, N), order="F") if return_ysim else None
    g_ms, b_ms = L.C.c_double(0.0), L.C.c_double(0.0)
    rc = L.lib().lqg_paths(n_t, m, N, L.ptr(Phi), n_t, L.ptr(xi), m, L.ptr(probs), probs.size,
                           L.ptr(ysim), n_t, L.ptr(mean), L.ptr(q), int(slab_rows), L.make_opts(mem=L.MEM_HOST),
                           L.C.byref(g_ms), L.C.byref(b_ms))
    L.check(rc, "lqg_paths")
    info = dict(gemm_ms=g_ms.value, bands_ms=b_ms.value)
    return (mean, q, ysim, info) if return_ysim else (mean, q, info)


def paths_bands(Phi_test, xi_sim, probs=(0.025, 0.975), slab_rows=None, return_ysim=False, xtest=None, ulist=None):
    """Band summaries of ysim = Phi.test %*% xi.sim without materialising ysim (config 5:
    10^5 x 10^6 doubles = 800 GB).  Phi.test (n_t x m) is processed in row slabs: each slab's
    sample paths are produced by the FP64 tensor-core GEMM into device memory, summarised by
    lqg_bands (exact type-7 quantiles + mean) and dropped.  xi.sim is uploaded once.
    With Phi_test = None and (xtest, ulist) given, each slab of the hat-basis matrix is built on
    the device from the test points (lqg_basis) instead of being computed on the host and
    uploaded (config 5: 800 MB).  xi_sim may also be a CUDA tensor of shape (N, m) (= m x N
    column-major) that is already on the device.
    Device buffers are torch tensors (plumbing); all compute goes through the C ABI.
    Returns (ymean[n_t], qtls[nprobs, n_t]) and timing info."""
    import torch
    xi_dev = xi_sim if isinstance(xi_sim, torch.Tensor) else None      # already on the device: (N, m) contiguous = m x N column-major
    xi = None if xi_dev is not None else np.asarray(xi_sim, float)
    if Phi_test is None:
        if xtest is None or ulist is None:
            raise ValueError("paths_bands needs Phi_test or (xtest, ulist)")
        if not isinstance(ulist, (list, tuple)):
            ulist = [ulist]
        D = len(ulist)
        xt = np.asarray(xtest, float).reshape(-1, D)
        m_k = np.ascontiguousarray([np.size(u) for u in ulist], dtype=np.int32)
        n_t, m = xt.shape[0], int(m_k.sum())
        Phi = None
    else:
        Phi = np.asarray(Phi_test, float)
        n_t, m = Phi.shape
    m2, N = (xi_dev.shape[1], xi_dev.shape[0]) if xi_dev is not None else xi.shape
    if m != m2:
        raise ValueError("non-conformable arguments")
    probs = np.ascontiguousarray(probs, dtype=np.float64)
    dev = torch.device("cuda", torch.cuda.current_device())
    st = torch.cuda.current_stream().cuda_stream
    if slab_rows is None:                                   # ysim slab + its transposed copy <= ~8 GB
        slab_rows = int(max(1, min(n_t, (4 << 30) // (8 * N))))
        if 128 <= slab_rows < n_t:                          # whole 128-row GEMM tiles (no partly filled row block)
            slab_rows -= slab_rows % 128
    if xi_dev is not None:
        if not (xi_dev.is_cuda and xi_dev.dtype == torch.float64 and xi_dev.is_contiguous()):
            raise ValueError("xi_sim on the device must be a contiguous float64 CUDA tensor of shape (N, m)")
        xi_d = xi_dev
    else:
        xi_d = torch.from_numpy(np.asfortranarray(xi).T.copy()).to(dev)           # column-major m x N
    y_d = torch.empty((N, slab_rows), dtype=torch.float64, device=dev)            # column-major slab x N
    mean_d = torch.empty(slab_rows, dtype=torch.float64, device=dev)
    q_d = torch.empty((slab_rows, probs.size), dtype=torch.float64, device=dev)   # column-major nprobs x slab
    probs_d = torch.from_numpy(probs).to(dev)
    if Phi is None:
        x_d = torch.from_numpy(np.asfortranarray(xt).T.copy()).to(dev)                # column-major n_t x D
        u_d = torch.from_numpy(np.concatenate([np.ravel(u) for u in ulist]).astype(np.float64)).to(dev)
        Phi_buf = torch.empty((m, slab_rows), dtype=torch.float64, device=dev)        # column-major slab x m
    mean = np.empty(n_t); q = np.empty((probs.size, n_t), order="F")
    o = L.make_opts(mem=L.MEM_DEVICE, stream=st)
    gemm_ms = bands_ms = 0.0
    ms = L.C.c_double(0.0)
    lib = L.lib()
    for r0 in range(0, n_t, slab_rows):
        r1 = min(n_t, r0 + slab_rows)
        nr = r1 - r0
        if Phi is None:
            L.check(lib.lqg_basis(nr, D, L.ptr(m_k), x_d.data_ptr() + 8 * r0, n_t, u_d.data_ptr(), o,
                                  Phi_buf.data_ptr(), nr), "lqg_basis")
            Phi_d = Phi_buf
        else:
            Phi_d = torch.from_numpy(np.asfortranarray(Phi[r0:r1]).T.copy()).to(dev)  # column-major nr x m
        L.check(lib.lqg_dgemm(nr, N, m, Phi_d.data_ptr(), nr, xi_d.data_ptr(), m, y_d.data_ptr(), nr, None, o, L.C.byref(ms)),
                "lqg_dgemm")
        gemm_ms += ms.value
        L.check(lib.lqg_bands(nr, N, y_d.data_ptr(), nr, probs_d.data_ptr(), probs.size, mean_d.data_ptr(), q_d.data_ptr(), o,
                              L.C.byref(ms)), "lqg_bands")
        bands_ms += ms.value
        mean[r0:r1] = mean_d[:nr].cpu().numpy()
        q[:, r0:r1] = q_d.reshape(-1)[: nr * probs.size].cpu().numpy().reshape(nr, probs.size).T
    info = dict(gemm_ms=gemm_ms, bands_ms=bands_ms, slab_rows=slab_rows,
                gemm_tflops=2.0 * n_t * m * N / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None)
    return mean, q, info


def simulate_pipeline(tmvPar, nsim, Lambda=None, Phi_test=None, xtest=None, ulist=None, probs=(0.025, 0.975),
                      control=None, comm=None, Lambda_pinv=None, return_eta=False, return_ysim=False, slab_rows=0):
    """The whole simulate tail in ONE C-ABI call (lqg_simulate): box HMC -> xi = Lambda^+ eta ->
    ysim = Phi.test xi -> mean / type-7 quantile bands, with eta, xi and ysim resident in HBM and
    only the requested results copied out (R/lineqGP.R:611-640; R/lineqAGP.R:558-575, 509-512;
    R/lineqGPutils.R:422-423).

    tmvPar: dict(mu, [map], Sigma, lb, ub) in eta space, as simulate.* builds it.  nsim = samples drawn
    by THIS rank; with `comm` (lineqgpr_b200.dist.Comm, one process per GPU) the job draws
    world * nsim samples, every rank ends up with the bands of all test points over all samples, and
    xi.sim holds this rank's columns.  Phi_test (n_t x m) or (xtest, ulist) to build the hat basis on
    the device; neither = stop after xi.  control keys as tmvrnorm.HMC ("n.chains", "burn.in", "seed",
    "chain.offset", "prepass", "max.restarts", "U").
    Returns a dict: xi.sim (m x nsim), ymean, qtls, [eta], [ysim], stats."""
    from .samplers import _as_par, EPSILON, chain_burn_in
    control = dict(control or {})
    control.setdefault("burn.in", 100)
    par = _as_par(tmvPar)
    mu, Sigma, lb, ub = par["mu"], par["Sigma"], par["lb"], par["ub"]
    d = mu.size
    if lb.size != d or ub.size != d:
        raise ValueError("The bounds have to be d-dimensionals")
    if np.any(lb >= ub):
        raise ValueError('The elements from "u" has to be greater than the elements from "l"')
    init = np.minimum(np.maximum(par["map"], lb + EPSILON), ub - EPSILON)
    nsim = int(nsim)
    slots = L.lib().lqg_hmc_box_slots(d) or 1024
    n_chains = max(1, min(int(control.get("n.chains", min(nsim, slots))), nsim))
    n_per = -(-nsim // n_chains)
    N_local = n_chains * n_per
    world = comm.world if comm is not None else 1
    N_total = N_local * world
    if Lambda is not None:
        Lambda = L.f64(np.asarray(Lambda, float)); m = Lambda.shape[1]
    elif Lambda_pinv is not None:
        Lambda_pinv = L.f64(np.asarray(Lambda_pinv, float)); m = Lambda_pinv.shape[0]
    else:
        m = d
    probs = np.ascontiguousarray(probs, dtype=np.float64)
    a = L.SimulateArgs()
    keep = []                                   # arrays that must outlive the call

    def P(x):
        if x is None:
            return None
        keep.append(x)
        return L.ptr(x)

    a.d, a.m = d, m
    a.mu, a.lb, a.ub, a.init = (P(np.ascontiguousarray(v, dtype=np.float64)) for v in (mu, lb, ub, init))
    a.Sigma = P(L.f64(Sigma))
    U = control.get("U")
    a.U = P(L.f64(U)) if U is not None else None
    a.Lambda, a.Lambda_pinv = P(Lambda), P(Lambda_pinv)
    n_t = 0
    if Phi_test is not None:
        Phi = L.f64(np.atleast_2d(np.asarray(Phi_test, float)))
        n_t = Phi.shape[0]
        if Phi.shape[1] != m:
            raise ValueError("non-conformable arguments")
        a.Phi, a.ldphi = P(Phi), n_t
    elif xtest is not None:
        if not isinstance(ulist, (list, tuple)):
            ulist = [ulist]
        D = len(ulist)
        xt = L.f64(np.asarray(xtest, float).reshape(-1, D))
        n_t = xt.shape[0]
        m_k = np.ascontiguousarray([np.size(u) for u in ulist], dtype=np.int32)
        knots = np.ascontiguousarray(np.concatenate([np.ravel(u) for u in ulist]), dtype=np.float64)
        a.basis_D, a.basis_mk, a.xtest, a.ldx, a.knots = D, P(m_k), P(xt), n_t, P(knots)
    a.n_t = n_t
    a.probs, a.nprobs = P(probs), probs.size
    a.n_per_chain = n_per
    a.burn_in = chain_burn_in(int(control["burn.in"]), n_chains * world, control)
    a.slab_rows = int(slab_rows)
    xi = control.get("out")
    if xi is None:
        xi = np.empty((m, N_local), order="F")
    eta = np.empty((d, N_local), order="F") if return_eta else None
    mean = np.empty(n_t) if n_t else None
    q = np.empty((probs.size, n_t), order="F") if n_t else None
    rows_per = -(-n_t // world) if n_t else 0
    rank = comm.rank if comm is not None else 0
    r0 = min(n_t, rank * rows_per); r1 = min(n_t, r0 + rows_per)
    ysim = np.empty((r1 - r0, N_total), order="F") if (return_ysim and n_t) else None
    a.eta_out, a.xi_out, a.ysim_out, a.ldy, a.mean, a.qtls = P(eta), P(xi), P(ysim), max(1, r1 - r0), P(mean), P(q)
    seed = control.get("seed")
    if seed is None:
        seed = int(np.random.randint(1, 10000000))                     # tmg:R/tmg.R:102
    from .samplers import chain_fork
    fork, fork_burn = chain_fork(n_chains, control)
    opts = L.make_opts(mem=L.MEM_HOST, seed=seed, n_chains=n_chains, max_restarts=int(control.get("max.restarts", 0)),
                       chain_offset=int(control.get("chain.offset", 0)), prepass=int(control.get("prepass", -1)),
                       hmc_search=int(control.get("hmc.search", 0)), fork=fork, fork_burn_in=fork_burn)
    st = L.SimulateStats()
    rc = L.lib().lqg_simulate(L.C.byref(a), opts, comm.handle if comm is not None else None, st)
    last_stats.clear(); last_stats.update(st.asdict(), sampler="HMC", n_chains=n_chains, n_per_chain=n_per, seed=seed,
                                          burn_in_per_chain=int(a.burn_in), fork=fork, fork_burn_in=fork_burn)
    if rc == 8:
        raise ValueError("Error: M must be positive definite.")        # tmg:R/tmg.R:17-20 (device whitening)
    L.check(rc, "lqg_simulate")
    out = {"xi.sim": xi[:, :nsim] if world == 1 else xi, "ymean": mean, "qtls": q, "stats": dict(last_stats),
           "rows": (r0, r1)}
    if return_eta:
        out["eta"] = eta[:, :nsim] if world == 1 else eta
    if ysim is not None:
        out["ysim"] = ysim
    return out


def lambda_pinv(Lambda):
    """The operator of qr.solve(Lambda, .) (R/lineqGP.R:639): inverse for a square Lambda, the
    least-squares pseudo-inverse for a tall full-column-rank Lambda (Q8).  O(d m^2), once."""
    Lambda = np.asarray(Lambda, float)
    if Lambda.shape[0] == Lambda.shape[1]:
        return np.linalg.inv(Lambda)
    Q, R = np.linalg.qr(Lambda)
    return np.linalg.solve(R, Q.T)


def _setup_mode(control):
    """control["setup"]: "host" (numpy/LAPACK, as R does it) or "device" (the C-ABI set-up stage:
    QP, factorizations, eta-space parameters and Lambda^+ on the GPU)."""
    mode = (control or {}).get("setup", "host")
    if mode not in ("host", "device"):
        raise ValueError(f"unknown setup {mode!r} (host | device)")
    return mode


def _pinv(Lambda, setup):
    if setup == "device":
        from .setup_stage import lambda_pinv_device
        return lambda_pinv_device(Lambda)
    return lambda_pinv(Lambda)


def simulate_lineqGP(model, nsim=1, seed=None, xtest=None, control=None):
    """simulate.lineqGP (R/lineqGP.R:601-643)."""
    setup = _setup_mode(control)
    t0 = time.perf_counter()
    pred = M.predict_lineqGP(model, xtest, setup=setup)
    predtime = time.perf_counter() - t0
    mdl = pred.pop("model")
    Lam = pred["Lambda"]
    if setup == "device":
        from . import setup_stage as st
        eta_mu, eta_map, Sigma_eta = st.eta_params(Lam, pred["mu"], pred["xi.map"], pred["Sigma"],
                                                   mdl["kernParam"]["nugget"])          # :611-614
    else:
        eta_mu = Lam @ pred["mu"]                                           # :611
        eta_map = Lam @ pred["xi.map"]                                      # :612
        Sigma_eta = Lam @ pred["Sigma"] @ Lam.T                             # :613
        Sigma_eta = Sigma_eta + mdl["kernParam"]["nugget"] * np.eye(Sigma_eta.shape[0])   # :614
    ctl = dict(mdl["localParam"]["samplingParams"])                     # :617
    ctl["mvec"] = mdl["localParam"]["mvec"]; ctl["constrType"] = mdl["constrType"]   # :618-619
    ctl.update(control or {})
    if setup == "device":
        ctl.setdefault("whiten", "device")
    if seed is not None:
        ctl.setdefault("seed", int(seed))                               # set.seed(seed) :624
    tmvPar = {"mu": eta_mu, "map": eta_map, "Sigma": Sigma_eta, "lb": pred["lb"], "ub": pred["ub"],
              "class": mdl["localParam"]["sampler"]}                    # :622-623
    extra = {}
    t0 = time.perf_counter()
    if tmvPar["class"] == "HMC" and ctl.get("pipeline", "device") == "device" and ctl.get("injected") is None:
        # :626, :639, :640 in one device-resident call; the bands plot.lineqGP draws come with it
        r = simulate_pipeline(tmvPar, nsim, Lambda=Lam, Phi_test=pred["Phi.test"], control=ctl,
                              probs=ctl.get("probs", (0.025, 0.975)), return_eta=True, return_ysim=True)
        simtime = time.perf_counter() - t0
        eta, xi_sim, ysim = r["eta"], r["xi.sim"], r["ysim"][:, :nsim]
        extra = {"ymean": r["ymean"], "qtls": r["qtls"]}
    else:
        eta = tmvrnorm(tmvPar, nsim, ctl)                               # :626
        simtime = time.perf_counter() - t0
        xi_sim = dgemm(_pinv(Lam, setup), eta)                          # :639
        ysim = dgemm(pred["Phi.test"], xi_sim)                          # :640
    return dict(cls="lineqGP", x=mdl["x"], y=mdl["y"], xtest=xtest, **{"Phi.test": pred["Phi.test"]},
                predtime=predtime, simtime=simtime, **{"xi.map": pred["xi.map"], "xi.sim": xi_sim},
                ysim=ysim, eta=eta, stats=dict(last_stats), **extra)


def simulate_lineqAGP(model, nsim=1, seed=None, xtest=None, control=None):
    """simulate.lineqAGP (R/lineqAGP.R:520-596)."""
    setup = _setup_mode(control)
    t0 = time.perf_counter()
    pred = M.predict_lineqAGP(model, xtest, setup=setup)
    predtime = time.perf_counter() - t0
    mdl = pred.pop("model")
    LamAll = mdl["lineqSys"]["LambdaAll"]
    if setup == "device":
        from . import setup_stage as st
        _, eta_map, Sigma_eta = st.eta_params(LamAll, pred["xiAll.map"], pred["xiAll.map"], pred["SigmaAll"],
                                              mdl["nugget"])            # :558-560
    else:
        eta_map = LamAll @ pred["xiAll.map"]                                # :558
        Sigma_eta = LamAll @ pred["SigmaAll"] @ LamAll.T                    # :559
        Sigma_eta = Sigma_eta + mdl["nugget"] * np.eye(Sigma_eta.shape[0])  # :560
    ctl = dict(mdl["localParam"]["samplingParams"])
    ctl["mvec"] = mdl["localParam"]["mtotal"]; ctl["constrType"] = mdl["constrType"]   # :564-565
    ctl.update(control or {})
    if setup == "device":
        ctl.setdefault("whiten", "device")
    if seed is not None:
        ctl.setdefault("seed", int(seed))
    tmvPar = {"mu": eta_map, "Sigma": Sigma_eta, "lb": mdl["lineqSys"]["lb"], "ub": mdl["lineqSys"]["ub"],
              "class": mdl["localParam"]["sampler"]}                    # :567-570 (mean := MAP, Q4)
    t0 = time.perf_counter()
    if tmvPar["class"] == "HMC" and ctl.get("pipeline", "device") == "device" and ctl.get("injected") is None:
        r = simulate_pipeline(tmvPar, nsim, Lambda=LamAll, control=ctl, return_eta=True)    # :573, :575 in one call
        simtime = time.perf_counter() - t0
        etaAll, xiAll_sim = r["eta"], r["xi.sim"]
    else:
        etaAll = tmvrnorm(tmvPar, nsim, ctl)                            # :573
        simtime = time.perf_counter() - t0
        xiAll_sim = dgemm(_pinv(LamAll, setup), etaAll)                 # :575
    return dict(cls="lineqAGP", x=mdl["x"], y=mdl["y"], xtest=xtest, **{"Phi.test": pred["Phi.test"]},
                predtime=predtime, simtime=simtime, muAll=pred["muAll"],
                **{"xiAll.map": pred["xiAll.map"], "xiAll.sim": xiAll_sim, "PhiAll.test": pred["PhiAll.test"]},
                eta=etaAll, stats=dict(last_stats))


def simulate(model, nsim=1, seed=None, xtest=None, control=None):
    """S3 dispatch of stats::simulate on the model class."""
    if model.get("cls") == "lineqGP":
        return simulate_lineqGP(model, nsim, seed, xtest, control)
    if model.get("cls") == "lineqAGP":
        return simulate_lineqAGP(model, nsim, seed, xtest, control)
    raise TypeError(f"no applicable method for 'simulate' applied to an object of class {model.get('cls')!r}")
