"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product package).

numpy / ctypes restatement of the R glue around the hot path of lineqGPR, each
function citing the reference lines it follows:

  R/...   = /root/reference/dev/lineqGPR/R/...
  tmg:... = file inside /root/reference/dependecies/tmg_0.3.tar.gz

The serial chain itself is oracle/tmg_hmc.cpp (liboracle_tmg.so); when
oracle/_ref/libtmg_ref.so exists (built by `make -C oracle ref` from the
unmodified reference HmcSampler.cpp) it can be driven through `ref_rtmg_canonical`.

Parity pinning: the reference ships no tests / golden vectors for this path
(SURVEY.md §4, §8c).  The chain restatement is pinned bit-for-bit against the
compiled reference source (tests/test_oracle.py::test_restatement_matches_reference_build)
and against analytic anchors (unconstrained limit, 1-D truncated normal).  The R
glue below (whitening, RSM, qr.solve, quantile) has no runnable reference in this
image (no R): for those functions parity is UNPINNED beyond analytic identities.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_dp = ctypes.POINTER(ctypes.c_double)
_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)


def _P(a):
    return a.ctypes.data_as(_dp)


def build(ref: bool = True) -> None:
    """Compile the C++ restatement and, when /root/reference is present, the reference build."""
    subprocess.run(["make", "-C", str(_HERE), "liboracle_tmg.so"], check=True, capture_output=True)
    if ref and Path("/root/reference/dependecies/tmg_0.3.tar.gz").exists():
        subprocess.run(["make", "-C", str(_HERE), "ref"], check=True, capture_output=True)


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        p = _HERE / "liboracle_tmg.so"
        if not p.exists():
            build(ref=False)
        _lib = ctypes.CDLL(str(p))
        _lib.oracle_rtmg.restype = ctypes.c_int
        _lib.oracle_rtmg.argtypes = [
            ctypes.c_int, ctypes.c_int, ctypes.c_int, _dp, ctypes.c_int, _dp, _dp,
            _dp, ctypes.c_int64, ctypes.c_int, ctypes.c_int64,
            _dp, _i32p, ctypes.c_int64, _dp, _i64p]
        _lib.oracle_normal_stream.restype = None
        _lib.oracle_normal_stream.argtypes = [ctypes.c_int, ctypes.c_int64, _dp]
        _lib.oracle_hit_time.restype = ctypes.c_double
        _lib.oracle_hit_time.argtypes = [ctypes.c_double] * 3
        _lib.oracle_hit_time_batch.restype = None
        _lib.oracle_hit_time_batch.argtypes = [ctypes.c_int64, _dp, _dp, _dp, _dp]
        _lib.oracle_rtmg_ext.restype = ctypes.c_int
        _lib.oracle_rtmg_ext.argtypes = [ctypes.c_int, ctypes.c_int, _dp, ctypes.c_int, _dp, _dp, _dp, ctypes.c_int64,
                                         ctypes.c_int64, _dp, _dp, _dp, _i64p]
    return _lib


def ref_lib():
    """The compiled UNMODIFIED reference (or None when it was never built)."""
    global _ref
    if _ref is None:
        p = _HERE / "_ref" / "libtmg_ref.so"
        if not p.exists():
            return None
        _ref = ctypes.CDLL(str(p))
        _ref.ref_rtmg.restype = ctypes.c_int
        _ref.ref_rtmg.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, _dp, ctypes.c_int,
                                  _dp, _dp, _dp]
    return _ref


# --------------------------------------------------------------------------- chain
def normal_stream(seed: int, n: int) -> np.ndarray:
    """The Gaussian draws tmg would make for `seed` (tmg:src/HmcSampler.h:50-52, .cpp:26,55)."""
    out = np.empty(int(n), dtype=np.float64)
    lib().oracle_normal_stream(int(seed), int(n), _P(out))
    return out


def hit_time(fa: float, fb: float, g: float) -> float:
    """tmg:src/HmcSampler.cpp:157-181 for one constraint."""
    return lib().oracle_hit_time(float(fa), float(fb), float(g))


def hit_time_batch(fa, fb, g):
    """oracle_hit_time over arrays (tmg:src/HmcSampler.cpp:157-181 per element)."""
    fa, fb, g = (np.ascontiguousarray(v, dtype=np.float64) for v in (fa, fb, g))
    out = np.empty_like(fa)
    lib().oracle_hit_time_batch(fa.size, _P(fa), _P(fb), _P(g), _P(out))
    return out


def rtmg_canonical_ext(n, initial, F, g, injected, unwhiten=None, shift=None, max_restarts=0):
    """The canonical chain in EXTENDED precision (oracle/tmg_hmc_ext.cpp), injected draws.  With
    `unwhiten` (= Ri) and `shift` (= mu) the result is Ri z + mu evaluated in extended precision
    (tmg:R/tmg.R:106).  Returns (samples n x d rounded to double, info)."""
    initial = np.ascontiguousarray(initial, dtype=np.float64)
    d = initial.size
    F = np.asfortranarray(np.asarray(F, dtype=np.float64).reshape(-1, d))
    g = np.ascontiguousarray(g, dtype=np.float64)
    inj = np.ascontiguousarray(injected, dtype=np.float64)
    out = np.zeros((n, d), order="F")
    stats = np.zeros(5, dtype=np.int64)
    uw = np.asfortranarray(unwhiten, dtype=np.float64) if unwhiten is not None else None
    sh = np.ascontiguousarray(shift, dtype=np.float64) if shift is not None else None
    rc = lib().oracle_rtmg_ext(int(n), d, _P(initial), F.shape[0], _P(F), _P(g), _P(inj), inj.size, int(max_restarts),
                               _P(uw) if uw is not None else None, _P(sh) if sh is not None else None, _P(out),
                               stats.ctypes.data_as(_i64p))
    return out, dict(rc=rc, bounces=int(stats[0]), vel_restarts=int(stats[1]), chk_restarts=int(stats[2]),
                     draws=int(stats[3]), hit_searches=int(stats[4]))


def rtmg_canonical(n, seed, initial, F, g, injected=None, faithful_copy=False,
                   max_restarts=0, trace_cap=0):
    """The `.Call("rtmg", n, seed, initial2, numlin, f2, g2, 0, NULL)` contract
    (tmg:R/tmg.R:104 -> tmg:src/rtmg.cpp:14-66).  Returns (samples n x d, info)."""
    initial = np.ascontiguousarray(initial, dtype=np.float64)
    d = initial.size
    F = np.asfortranarray(np.asarray(F, dtype=np.float64).reshape(-1, d))
    c = F.shape[0]
    g = np.ascontiguousarray(g, dtype=np.float64)
    assert g.size == c
    out = np.zeros((n, d), order="F")
    stats = np.zeros(6, dtype=np.int64)
    inj = None
    n_inj = 0
    if injected is not None:
        inj = np.ascontiguousarray(injected, dtype=np.float64)
        n_inj = inj.size
    tt = np.zeros(max(trace_cap, 1))
    tc = np.zeros(max(trace_cap, 1), dtype=np.int32)
    rc = lib().oracle_rtmg(int(n), int(seed), d, _P(initial), c, _P(F), _P(g),
                           _P(inj) if inj is not None else None, n_inj,
                           int(bool(faithful_copy)), int(max_restarts),
                           _P(tt), tc.ctypes.data_as(_i32p), int(trace_cap),
                           _P(out), stats.ctypes.data_as(_i64p))
    info = dict(rc=rc, bounces=int(stats[0]), vel_restarts=int(stats[1]),
                chk_restarts=int(stats[2]), draws=int(stats[3]), hit_searches=int(stats[4]),
                n_trace=int(stats[5]), trace_t=tt[:min(trace_cap, stats[5])],
                trace_cn=tc[:min(trace_cap, stats[5])])
    return out, info


def ref_rtmg_canonical(n, seed, initial, F, g):
    """Same contract, executed by the compiled reference source (oracle/_ref)."""
    r = ref_lib()
    if r is None:
        raise RuntimeError("oracle/_ref/libtmg_ref.so not built (make -C oracle ref)")
    initial = np.ascontiguousarray(initial, dtype=np.float64)
    d = initial.size
    F = np.asfortranarray(np.asarray(F, dtype=np.float64).reshape(-1, d))
    g = np.ascontiguousarray(g, dtype=np.float64)
    out = np.zeros((n, d), order="F")
    r.ref_rtmg(int(n), int(seed), d, _P(initial), F.shape[0], _P(F), _P(g), _P(out))
    return out


# --------------------------------------------------------------------------- R glue
def bounds2lineqSys_oneside(d, l, u, A=None, rmInf=True):
    """R/lineqGPutils.R:353-390, lineqSysType='oneside': M=[A;-A], g=[-l;u], drop g==Inf."""
    l = np.asarray(l, dtype=np.float64).ravel()
    u = np.asarray(u, dtype=np.float64).ravel()
    if l.size not in (1, d) or u.size not in (1, d):
        raise ValueError("The bounds have to be d-dimensionals")            # :355-356
    if A is None:
        A = np.eye(d)
    if A.shape[0] != d:
        raise ValueError("nrow(A) has to be equal to d")                     # :357-358
    if np.any(l >= u):
        raise ValueError('The elements from "u" has to be greater than the elements from "l"')
    if l.size == 1:
        l = np.repeat(l, d)
    if u.size == 1:
        u = np.repeat(u, d)
    g = np.concatenate([-l, u])                                              # :375
    M = np.vstack([A, -A])                                                   # :376
    if rmInf:
        keep = ~(g == np.inf)                                                # :377-381
        g, M = g[keep], M[keep]
    return M, g


def whiten(M, r, initial, f, g):
    """tmg:R/tmg.R:15-28,44-51.  Returns dict(R, Ri, Mir, initial2, f2, g2)."""
    M = (M + M.T) / 2                                                        # :15
    eigs = np.linalg.eigvalsh(M)                                             # :16
    if np.any(eigs <= 0):
        raise ValueError("M must be positive definite.")                     # :17-20
    R = np.linalg.cholesky(M).T                                              # :25 (upper, R'R = M)
    Mir = np.linalg.solve(M, r)                                              # :26
    Ri = np.linalg.inv(R)                                                    # :27
    initial2 = R @ initial - Ri.T @ r                                        # :28
    if np.any(f @ initial + g <= 0):
        raise ValueError("initial point violates linear constraints.")       # :44-47
    f2 = f @ Ri                                                              # :50
    g2 = f @ Mir + g                                                         # :51
    return dict(R=R, Ri=Ri, Mir=Mir, initial2=initial2, f2=f2, g2=g2)


def rtmg(n, M, r, initial, f, g, burn_in=30, seed=1, injected=None, faithful_copy=False,
         use_reference_build=False):
    """tmg:R/tmg.R:1-108 (linear constraints only).  Returns (n x d samples, info).
    `seed` stands in for R's sample(1:1e7, 1) (tmg:R/tmg.R:102)."""
    w = whiten(np.asarray(M, float), np.asarray(r, float), np.asarray(initial, float),
               np.asarray(f, float), np.asarray(g, float))
    if use_reference_build:
        z = ref_rtmg_canonical(n + burn_in, seed, w["initial2"], w["f2"], w["g2"])
        info = {}
    else:
        z, info = rtmg_canonical(n + burn_in, seed, w["initial2"], w["f2"], w["g2"],
                                 injected=injected, faithful_copy=faithful_copy)
    z = z[burn_in:burn_in + n]                                               # :105
    x = z @ w["Ri"].T + w["Mir"][None, :]                                    # :106
    info.update(w)
    info["z"] = z
    return x, info


def hmc_setup(mu, Sigma, lb, ub, map_=None):
    """The part of tmvrnorm.HMC before the rtmg call (R/lineqGPsamplers.R:99-139):
    returns (H, r, init, f, g)."""
    mu = np.asarray(mu, float).ravel()
    d = mu.size
    if map_ is None:
        map_ = mu                                                            # :107-108
    Mm, g = bounds2lineqSys_oneside(d, lb, ub, np.eye(d))                    # :128-132 (mvec := d, Q3)
    eps = 1e-9
    init = np.minimum(np.maximum(map_, np.asarray(lb, float) + eps), np.asarray(ub, float) - eps)  # :133-135
    H = np.linalg.inv(np.asarray(Sigma, float))                              # :139 solve(Sigma)
    r = mu @ H                                                               # :140
    return H, r, init, Mm, g


def tmvrnorm_HMC(obj, nsim, control=None, seed=1, injected=None, **kw):
    """R/lineqGPsamplers.R:99-143.  Returns (d x nsim, info)."""
    control = dict(control or {})
    burn = int(control.get("burn.in", 100))                                  # :100-101
    H, r, init, f, g = hmc_setup(obj["mu"], obj["Sigma"], obj["lb"], obj["ub"], obj.get("map"))
    x, info = rtmg(nsim, H, r, init, f, g, burn_in=burn, seed=seed, injected=injected, **kw)
    info["init"] = init
    return x.T, info                                                         # :142


def mvrnorm_factor(Sigma):
    """MASS::mvrnorm's factor (source not vendored; documented algorithm):
    eS = eigen(Sigma, symmetric=TRUE); X = mu + eS$vectors %*% diag(sqrt(pmax(ev,0))) %*% z."""
    ev, V = np.linalg.eigh(np.asarray(Sigma, float))
    ev, V = ev[::-1], V[:, ::-1]                                             # R orders decreasing
    return V * np.sqrt(np.maximum(ev, 0.0))[None, :]


def tmvrnorm_RSM(obj, nsim, z_stream, u_stream, factor=None, max_proposals=None):
    """R/lineqGPsamplers.R:36-61 with injected draws.
    z_stream: d x P normals, one column per proposal (rnorm(d) of each mvrnorm call);
    u_stream: uniforms consumed one per IN-BOX proposal (:55).
    Returns (d x nsim, info) with info['accepted_idx'] = proposal indices accepted,
    info['inbox'] flags and info['n_proposals'] consumed."""
    mu = np.asarray(obj["mu"], float).ravel()
    map_ = np.asarray(obj.get("map", mu), float).ravel()                     # :39-40
    Sigma = np.asarray(obj["Sigma"], float)
    lb = np.asarray(obj["lb"], float).ravel()
    ub = np.asarray(obj["ub"], float).ravel()
    d = map_.size
    invSigma = _chol2inv(Sigma)                                              # :42
    w = invSigma @ map_                                                      # :43
    if factor is None:
        factor = mvrnorm_factor(Sigma)
    P = z_stream.shape[1] if max_proposals is None else min(max_proposals, z_stream.shape[1])
    xi = np.zeros((d, nsim))
    acc, inbox = [], np.zeros(P, dtype=bool)
    p = 0
    ui = 0
    k = 0
    while k < nsim and p < P:
        x = map_ + factor @ z_stream[:, p]                                   # :50 / :52
        if np.all((x >= lb) & (x <= ub)):                                    # :51
            inbox[p] = True
            t = np.exp((map_ - x) @ w)                                       # :54
            if ui >= len(u_stream):
                break
            u = u_stream[ui]; ui += 1                                        # :55
            if u <= t:                                                       # :56
                xi[:, k] = x                                                 # :58
                acc.append(p)
                k += 1
        p += 1
    return xi[:, :k], dict(accepted_idx=np.array(acc, dtype=np.int64), inbox=inbox[:p],
                           n_proposals=p, n_uniforms=ui, w=w, factor=factor)


def _chol2inv(S):
    """chol2inv(chol(S)) of R (R/lineqGPsamplers.R:42)."""
    L = np.linalg.cholesky(S)
    Li = np.linalg.inv(L)
    return Li.T @ Li


def qr_solve(Lambda, eta):
    """qr.solve(Lambda, eta) (R/lineqGP.R:639, R/lineqAGP.R:575): exact solve for square
    Lambda, least squares for tall full-rank Lambda (Q8)."""
    Lambda = np.asarray(Lambda, float)
    if Lambda.shape[0] == Lambda.shape[1]:
        return np.linalg.solve(Lambda, eta)
    return np.linalg.lstsq(Lambda, eta, rcond=None)[0]


def simulate_tail(Lambda, Phi_test, eta):
    """R/lineqGP.R:639-640: xi.sim = qr.solve(Lambda, eta); ysim = Phi.test %*% xi.sim."""
    xi = qr_solve(Lambda, eta)
    return xi, np.asarray(Phi_test, float) @ xi


def quantile7(x, probs):
    """stats::quantile default (type 7) along the last axis; used by the band summaries
    (R/lineqGPutils.R:422, :296).  h=(n-1)p; lo=floor(h); x[lo] + (h-lo)(x[lo+1]-x[lo])."""
    x = np.sort(np.asarray(x, float), axis=-1)
    n = x.shape[-1]
    out = []
    for p in np.atleast_1d(probs):
        fuzz = 4 * np.finfo(float).eps
        nppm = (n - 1) * p
        lo = int(np.floor(nppm + fuzz))                   # R: j <- floor(nppm + fuzz)
        hi = min(lo + 1, n - 1)
        h = nppm - lo
        if abs(h) < fuzz:
            h = 0.0
        xlo, xhi = x[..., lo], x[..., hi]
        q = np.where((h > 0) & (h < 1) & (xlo != xhi), (1 - h) * xlo + h * xhi, xlo)
        if h == 1:
            q = xhi
        out.append(q)
    return np.stack(out, axis=0)


def bands(ysim, probs=(0.025, 0.975)):
    """R/lineqGPutils.R:422-423: qtls = apply(ysim,1,quantile,probs); ymean = apply(ysim,1,mean)."""
    ysim = np.asarray(ysim, float)
    return ysim.mean(axis=1), quantile7(ysim, probs)
