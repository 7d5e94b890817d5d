"""`tmvrnorm` and its samplers: the reference's operator interface for the hot path, same names,
argument meaning and error behaviour, with the bodies replaced by calls into the C ABI
(include/lineqgpr_b200.h).  Mirrors

  tmvrnorm generic      R/lineqGPutils.R:57-58     (S3 dispatch on class(object))
  tmvrnorm.HMC          R/lineqGPsamplers.R:99-143 (+ tmg::rtmg's R wrapper, tmg:R/tmg.R:1-108)
  tmvrnorm.RSM          R/lineqGPsamplers.R:36-61
  tmvrnorm.ExpT         R/lineqGPsamplers.R:179-185

`object` is the R list `list(mu, [map], Sigma, lb, ub)` as a dict; its S3 class is the key
"class" (or the `method=` argument).  The result is a d x nsim array, samples in columns.

What stays on the host, exactly where the reference has it on the host: the one-off O(d^3)
whitening algebra (solve, chol, triangular inverse).  Everything per-sample runs on the GPU.
Extensions (all optional keys of `control`): "n.chains" (independent chains; 1 reproduces the
reference's single serial chain; with several chains each one runs at least 100 burn-in
transitions, see chain_burn_in), "burn.in.exact", "fork" / "fork.burn.in" (families of chains sharing a root chain's burn-in, see
chain_fork), "seed", "prepass", "injected", "max.restarts", "hmc.search" (lqg_opts.hmc_search),
"chain.offset", "U" (a precomputed factor), "whiten" ("device" (default) | "ul" | "tmg", see whiten_box).
"""
from __future__ import annotations

import numpy as np

from . import _lib as L

EPSILON = 1e-9       # R/lineqGPsamplers.R:127


class TmvStats(dict):
    """Counters of the last sampler call (bounces, restarts, timings ...)."""


last_stats = TmvStats()


def _as_par(object):
    tmvPar = dict(object)
    mu = np.asarray(tmvPar["mu"], float).ravel()
    if "map" not in tmvPar or tmvPar["map"] is None:
        tmvPar["map"] = mu                                             # R/lineqGPsamplers.R:39-40, :107-108
    tmvPar["mu"] = mu
    tmvPar["map"] = np.asarray(tmvPar["map"], float).ravel()
    tmvPar["Sigma"] = np.asarray(tmvPar["Sigma"], float)
    d = mu.size
    for k in ("lb", "ub"):
        v = np.asarray(tmvPar[k], float).ravel()
        tmvPar[k] = np.repeat(v, d) if v.size == 1 else v
    return tmvPar


def whiten_box(mu, Sigma, method="ul"):
    """The factor tmg::rtmg whitens with for tmvrnorm.HMC's call (R/lineqGPsamplers.R:139-141 ->
    tmg:R/tmg.R:15-28): H = solve(Sigma); M = (H+H')/2; positive-definiteness check;
    R = chol(M); Ri = solve(R) — an upper-triangular Ri with Ri Ri' = M^-1 = Sigma.

    method="tmg": that literal sequence (inverse, Cholesky, triangular inverse; ~8/3 d^3 flop and
                  one inversion of an ill-conditioned Sigma) — used by the parity tests.
    method="device": method "ul" computed on the GPU (lqg_whiten_box, blocked factorisation on
                  the FP64 tensor cores).
    method="ul" : the same matrix obtained directly: the upper-triangular factor with positive
                  diagonal and U U' = Sigma is unique, and equals J chol(J Sigma J) J with J the
                  exchange matrix (a "reverse Cholesky"): one Cholesky, d^3/3 flop, no inversion.
    Both raise tmg's error when Sigma (equivalently M) is not positive definite."""
    Sigma = np.asarray(Sigma, float)
    if method == "device":
        from .setup_stage import whiten_box_device
        return whiten_box_device(Sigma)
    if method == "ul":
        S = (Sigma + Sigma.T) / 2
        try:
            Lf = np.linalg.cholesky(S[::-1, ::-1])
        except np.linalg.LinAlgError:
            raise ValueError("Error: M must be positive definite.")    # tmg:R/tmg.R:17-20
        return np.ascontiguousarray(Lf[::-1, ::-1])
    if method != "tmg":
        raise ValueError(f"unknown whitening method {method!r}")
    H = np.linalg.inv(Sigma)                                           # solve(Sigma)
    M = (H + H.T) / 2                                                  # tmg:R/tmg.R:15
    try:
        Rl = np.linalg.cholesky(M)                                     # R = chol(M) = Rl'
    except np.linalg.LinAlgError:
        raise ValueError("Error: M must be positive definite.")        # tmg:R/tmg.R:17-20
    import scipy.linalg as sla
    Ri = sla.solve_triangular(Rl.T, np.eye(M.shape[0]), lower=False)   # Ri = solve(R)
    return np.triu(Ri)


MULTI_CHAIN_BURN_IN = 100   # = the burn.in default of tmvrnorm.HMC (R/lineqGPsamplers.R:100-101)


def chain_burn_in(burn_in, n_chains, control=None):
    """Burn-in per chain.  The reference runs ONE chain of nsim + burn.in transitions from the
    clamped MAP, and simulate.* hands it burn.in = 1 (R/lineqGP.R:264,617): its samples are spread
    along a long chain.  With n.chains > 1 every chain starts at that same point, so a burn-in of 1
    would make every column the second transition after the MAP — not the reference's chain
    distribution under active constraints.  Several chains therefore never run fewer than
    MULTI_CHAIN_BURN_IN transitions each before emitting (control["burn.in.exact"] = True, or an
    injected draw stream, keeps the caller's value: parity tests); n.chains = 1 is the reference
    chain and keeps burn.in as given."""
    control = control or {}
    if n_chains > 1 and not control.get("burn.in.exact") and control.get("injected") is None:
        return max(int(burn_in), MULTI_CHAIN_BURN_IN)
    return int(burn_in)


FORK_DEFAULT, FORK_BURN_IN = 6, 5


def chain_fork(n_chains, control=None):
    """(fork, fork_burn_in) of lqg_opts.  With many chains, families of 6 share one root chain's burn-in
    (the root runs the burn-in from the clamped MAP; its 6 chains start from the state it reached, run
    5 more transitions on their own streams, then emit): a sixth of the burn-in work, and every emitted
    column is still burn.in + 5 transitions away from the start.  control["fork"] = 1 gives fully
    independent chains (each with its own burn-in); injected draws never fork."""
    control = control or {}
    if control.get("injected") is not None:
        return 0, 0
    # (the default must not depend on how many chains THIS call runs: a sharded job has to pick the same
    #  families as the whole job; a lone chain without an offset is the reference's single serial chain)
    fork = int(control.get("fork", FORK_DEFAULT if (n_chains > 1 or "chain.offset" in control) else 1))
    return (fork, int(control.get("fork.burn.in", FORK_BURN_IN))) if fork > 1 else (0, 0)


def tmvrnorm_HMC(object, nsim, control=None, **kw):
    """tmvrnorm.HMC (R/lineqGPsamplers.R:99-143) on the GPU: box-form exact HMC, lqg_hmc_box."""
    control = dict(control or {})
    control.setdefault("burn.in", 100)                                 # :100-101
    tmvPar = _as_par(object)
    mu, Sigma, lb, ub = tmvPar["mu"], tmvPar["Sigma"], tmvPar["lb"], tmvPar["ub"]
    d = mu.size
    control.setdefault("mvec", d)                                      # :102-103 (treated as length(mu): Q3)
    if lb.size != d or ub.size != d:
        raise ValueError("The bounds have to be d-dimensionals")       # R/lineqGPutils.R:355-356
    if np.any(lb >= ub):
        raise ValueError('The elements from "u" has to be greater than the elements from "l"')
    init = np.minimum(np.maximum(tmvPar["map"], lb + EPSILON), ub - EPSILON)   # :133-135
    nsim = int(nsim)
    # default: one chain per resident CTA slot of the device (every wave full, least burn-in work)
    slots = L.lib().lqg_hmc_box_slots(d) or 1024
    n_chains = int(control.get("n.chains", min(nsim, slots)))
    n_chains = max(1, min(n_chains, nsim))
    n_per = -(-nsim // n_chains)
    U = control.get("U")
    if U is None and control.get("whiten", "device") != "device":
        U = whiten_box(mu, Sigma, control["whiten"])               # default "device": lqg_hmc_box factors Sigma itself
    seed = control.get("seed")
    if seed is None:
        seed = int(np.random.randint(1, 10000000))                     # tmg:R/tmg.R:102
    injected = control.get("injected")
    n_inj = 0
    if injected is not None:
        injected = np.ascontiguousarray(injected, dtype=np.float64).reshape(n_chains, -1)
        n_inj = injected.shape[1]
    fork, fork_burn = chain_fork(n_chains, control)
    opts = L.make_opts(mem=L.MEM_HOST, seed=seed, n_chains=n_chains,
                       max_restarts=int(control.get("max.restarts", 0)),
                       chain_offset=int(control.get("chain.offset", 0)),
                       injected=injected, n_injected=n_inj, prepass=int(control.get("prepass", -1)),
                       hmc_search=int(control.get("hmc.search", 0)), fork=fork, fork_burn_in=fork_burn)
    Uf, Sf = (L.f64(U) if U is not None else None), L.f64(Sigma)
    mu_c, lb_c, ub_c, init_c = (np.ascontiguousarray(v, dtype=np.float64) for v in (mu, lb, ub, init))
    out = control.get("out")                                            # optional caller-owned d x (n_chains*n_per) F-array
    if out is None:                                                    # (e.g. _lib.empty_result(): page-locked, reusable)
        out = np.empty((d, n_chains * n_per), order="F")
    elif not (isinstance(out, np.ndarray) and out.dtype == np.float64 and out.flags.f_contiguous
              and out.flags.writeable and out.shape == (d, n_chains * n_per)):
        raise ValueError(f'control["out"] must be a writable Fortran-ordered float64 array of shape ({d}, {n_chains * n_per})')
    burn = chain_burn_in(int(control["burn.in"]), n_chains, control)
    stats = L.HmcStats()
    rc = L.lib().lqg_hmc_box(d, L.ptr(mu_c), L.ptr(Uf), 1, L.ptr(Sf), L.ptr(lb_c), L.ptr(ub_c),
                             L.ptr(init_c), 0, n_per, burn, opts, L.ptr(out), None, stats)
    last_stats.clear(); last_stats.update(stats.asdict(), sampler="HMC", n_chains=n_chains, n_per_chain=n_per, seed=seed,
                                          burn_in_per_chain=burn, fork=fork, fork_burn_in=fork_burn)
    if rc == 8:
        raise ValueError("Error: M must be positive definite.")        # tmg:R/tmg.R:17-20 (device whitening)
    L.check(rc, "tmvrnorm.HMC")
    return out[:, :nsim]                                               # :142  t(xi): d x nsim


def mvrnorm_factor(Sigma):
    """MASS::mvrnorm's factor eS$vectors %*% diag(sqrt(pmax(ev, 0))) (the reference recomputes
    eigen(Sigma) at every proposal, R/lineqGPsamplers.R:50; it is hoisted here)."""
    ev, V = np.linalg.eigh(np.asarray(Sigma, float))
    ev, V = ev[::-1], V[:, ::-1]
    return V * np.sqrt(np.maximum(ev, 0.0))[None, :]


def tmvrnorm_RSM(object, nsim, control=None, **kw):
    """tmvrnorm.RSM (R/lineqGPsamplers.R:36-61) on the GPU: batched proposals, lqg_rsm."""
    control = dict(control or {})
    tmvPar = _as_par(object)
    map_, Sigma, lb, ub = tmvPar["map"], tmvPar["Sigma"], tmvPar["lb"], tmvPar["ub"]
    d = map_.size
    Lc = np.linalg.cholesky(Sigma)                                     # chol2inv(chol(Sigma)) :42
    Li = np.linalg.inv(Lc)
    w = (Li.T @ Li) @ map_                                             # :43
    factor = control.get("factor")
    if factor is None:
        factor = mvrnorm_factor(Sigma)
    seed = control.get("seed")
    if seed is None:
        seed = int(np.random.randint(1, 2 ** 31 - 1))
    inj_z, inj_u = control.get("injected"), control.get("injected.u")
    n_z = n_u = 0
    if inj_z is not None:
        inj_z = L.f64(np.asarray(inj_z, float).reshape(d, -1)); n_z = inj_z.shape[1]
    if inj_u is not None:
        inj_u = np.ascontiguousarray(inj_u, dtype=np.float64); n_u = inj_u.size
    opts = L.make_opts(mem=L.MEM_HOST, seed=seed, injected=inj_z, n_injected=n_z, injected_u=inj_u, n_injected_u=n_u)
    nsim = int(nsim)
    out = np.empty((d, nsim), order="F")
    stats = L.RsmStats()
    ff = L.f64(factor)
    map_c, w_c, lb_c, ub_c = (np.ascontiguousarray(v, dtype=np.float64) for v in (map_, w, lb, ub))
    rc = L.lib().lqg_rsm(d, L.ptr(map_c), L.ptr(ff), L.ptr(w_c), L.ptr(lb_c), L.ptr(ub_c), nsim,
                         int(control.get("max.proposals", 0)), opts, L.ptr(out), stats)
    last_stats.clear(); last_stats.update(stats.asdict(), sampler="RSM", seed=seed)
    L.check(rc, "tmvrnorm.RSM")
    return out


def tmvrnorm_ExpT(object, nsim, control=None, **kw):
    """tmvrnorm.ExpT (R/lineqGPsamplers.R:179-185):
        xi <- mu + TruncatedNormal::mvrandn(lb - mu, ub - mu, Sigma, nsim)
    Set-up on the host as in the R package (lineqgpr_b200/expt.py: permuted Cholesky, minimax tilt),
    proposals / accept / compaction on the GPU (lqg_expt), back-transformation Lfull[order,] Z on
    the FP64 tensor cores (lqg_dgemm).  Samples are i.i.d. and exact; TruncatedNormal's source is
    not in the reference tree, so parity with R is unpinned (Botev 2017 restated)."""
    from . import expt
    control = dict(control or {})
    tmvPar = _as_par(object)
    mu, Sigma, lb, ub = tmvPar["mu"], tmvPar["Sigma"], tmvPar["lb"], tmvPar["ub"]
    d = mu.size
    nsim = int(nsim)
    su = control.get("setup")
    if su is None or isinstance(su, str):              # "device": lqg_expt_setup; default: the host restatement
        su = expt.setup_device(lb - mu, ub - mu, Sigma) if su == "device" else expt.setup(lb - mu, ub - mu, Sigma)
    seed = control.get("seed")
    if seed is None:
        seed = int(np.random.randint(1, 2 ** 31 - 1))
    opts = L.make_opts(mem=L.MEM_HOST, seed=seed)
    L0 = L.f64(su["L0"])
    l_c, u_c, m_c = (np.ascontiguousarray(v, dtype=np.float64) for v in (su["l"], su["u"], su["mu"]))
    Z = np.empty((d, nsim), order="F")
    stats = L.ExptStats()
    rc = L.lib().lqg_expt(d, L.ptr(L0), L.ptr(l_c), L.ptr(u_c), L.ptr(m_c), float(su["psistar"]), nsim,
                          int(control.get("max.proposals", 0)), opts, L.ptr(Z), stats)
    last_stats.clear(); last_stats.update(stats.asdict(), sampler="ExpT", seed=seed, psistar=float(su["psistar"]))
    L.check(rc, "tmvrnorm.ExpT")
    if control.get("return.Z"):
        return Z
    A = L.f64(su["Lfull"][su["order"], :])                              # rv[order, ] = (Lfull %*% Z)[order, ]
    C_ = np.empty((d, nsim), order="F")
    ms = L.C.c_double(0.0)
    rc = L.lib().lqg_dgemm(d, nsim, d, L.ptr(A), d, L.ptr(Z), d, L.ptr(C_), d, None, L.make_opts(mem=L.MEM_HOST), L.C.byref(ms))
    L.check(rc, "tmvrnorm.ExpT (L Z)")
    return C_ + mu[:, None]


_METHODS = {"HMC": tmvrnorm_HMC, "RSM": tmvrnorm_RSM, "ExpT": tmvrnorm_ExpT}


def tmvrnorm(object, nsim, control=None, method=None, **kw):
    """tmvrnorm <- function(object, nsim, ...) UseMethod("tmvrnorm")   (R/lineqGPutils.R:57-58)."""
    cls = method or object.get("class")
    if cls not in _METHODS:
        raise TypeError(f'no applicable method for \'tmvrnorm\' applied to an object of class "{cls}"')
    return _METHODS[cls](object, nsim, control, **kw)


# --------------------------------------------------------------------------- canonical-frame access
def rtmg_canonical(n, seed, initial, F, g, n_chains=1, burn_in=0, injected=None, max_restarts=0, mode=0):
    """The reference's native contract, .Call("rtmg", n, seed, initial, numlin, F, g, 0, NULL)
    (tmg:R/tmg.R:104), served by lqg_hmc_canonical.  Returns (d x (n_chains*n) samples in the
    canonical frame, stats).  mode: 0 auto, 1 per-chain kernel, 2 lock-step GEMM rounds."""
    initial = np.asarray(initial, float)
    d = initial.shape[0]
    F = L.f64(np.asarray(F, float).reshape(-1, d))
    g = np.ascontiguousarray(g, dtype=np.float64)
    c = F.shape[0]
    n_inj = 0
    if injected is not None:
        injected = np.ascontiguousarray(injected, dtype=np.float64).reshape(n_chains, -1)
        n_inj = injected.shape[1]
    ld = 0 if initial.ndim == 1 else d
    init_c = np.ascontiguousarray(initial.T if initial.ndim == 2 else initial, dtype=np.float64)
    opts = L.make_opts(mem=L.MEM_HOST, seed=seed, n_chains=n_chains, max_restarts=max_restarts,
                       injected=injected, n_injected=n_inj, canon_mode=mode)
    out = np.empty((d, n_chains * n), order="F")
    stats = L.HmcStats()
    rc = L.lib().lqg_hmc_canonical(d, c, L.ptr(F), L.ptr(g), L.ptr(init_c), ld, int(n), int(burn_in),
                                   opts, L.ptr(out), None, stats)
    L.check(rc, "rtmg")
    return out, stats.asdict()


def rtmg_canonical_trace(n, initial, F, g, injected, n_chains=1, trace_cap=256):
    """Parity tooling: lqg_hmc_canonical_trace.  Returns (samples d x (n_chains*n), trace_t, trace_cn,
    trace_n) with the per-chain sequence of wall hits (t, cn) of tmg's _getNextLinearHitTime."""
    initial = np.asarray(initial, float)
    d = initial.shape[0]
    F = L.f64(np.asarray(F, float).reshape(-1, d))
    g = np.ascontiguousarray(g, dtype=np.float64)
    injected = np.ascontiguousarray(injected, dtype=np.float64).reshape(n_chains, -1)
    ld = 0 if initial.ndim == 1 else d
    init_c = np.ascontiguousarray(initial.T if initial.ndim == 2 else initial, dtype=np.float64)
    opts = L.make_opts(mem=L.MEM_HOST, n_chains=n_chains, injected=injected, n_injected=injected.shape[1])
    out = np.empty((d, n_chains * n), order="F")
    tt = np.zeros((n_chains, trace_cap)); tc = np.zeros((n_chains, trace_cap), dtype=np.int32)
    tn = np.zeros(n_chains, dtype=np.int32)
    rc = L.lib().lqg_hmc_canonical_trace(d, F.shape[0], L.ptr(F), L.ptr(g), L.ptr(init_c), ld, int(n), opts, L.ptr(out),
                                         int(trace_cap), L.ptr(tt), L.ptr(tc), L.ptr(tn))
    L.check(rc, "rtmg (trace)")
    return out, tt, tc, tn
