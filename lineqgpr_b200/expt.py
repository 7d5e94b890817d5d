"""Host set-up of the minimax exponential-tilting sampler behind tmvrnorm.ExpT.

tmvrnorm.ExpT (R/lineqGPsamplers.R:179-185) is
    mu + TruncatedNormal::mvrandn(lb - mu, ub - mu, Sigma, nsim).
TruncatedNormal (CRAN; imported unversioned at DESCRIPTION:12) is NOT vendored in the reference
tree, so this restates the published algorithm it implements:
    Z. I. Botev (2017), "The normal law under linear restrictions: simulation and estimation via
    minimax tilting", JRSS-B 79(1):125-148 — Algorithm 1 (permuted Cholesky of Gibson, Glasbey &
    Elston), the tilting parameters (x*, mu*) as the root of grad psi (eq. 11), sequential
    importance proposals and the accept test  -log U > psi* - logpr.
Parity is UNPINNED against R (no source, no fixtures); checks are distributional.

What runs where: exactly as in the R package, the O(d^3) set-up below (permuted Cholesky, Newton
solve) is host code; the per-sample work (proposal sweep, weights, accept, compaction, L Z) is
on the GPU (csrc/lqg_expt.cu, lqg_dgemm)."""
from __future__ import annotations

import numpy as np
from scipy.special import log_ndtr, ndtr

EPS = np.finfo(float).eps


def lnNpr(a, b):
    """log of the standard normal mass of [a, b], stable in both tails."""
    a = np.asarray(a, float); b = np.asarray(b, float)
    p = np.zeros(np.broadcast(a, b).shape)
    a, b = np.broadcast_to(a, p.shape), np.broadcast_to(b, p.shape)
    up = a > 0                                  # 0 < a < b: work with upper tails
    if up.any():
        pa = log_ndtr(-a[up]); pb = log_ndtr(-b[up])
        p[up] = pa + np.log1p(-np.exp(pb - pa))
    lo = b < 0                                  # a < b < 0: lower tails
    if lo.any():
        pa = log_ndtr(a[lo]); pb = log_ndtr(b[lo])
        p[lo] = pb + np.log1p(-np.exp(pa - pb))
    mid = ~up & ~lo                             # a <= 0 <= b
    if mid.any():
        p[mid] = np.log1p(-ndtr(a[mid]) - ndtr(-b[mid]))
    return p


def cholperm(Sig, l, u):
    """Cholesky factor with the variable reordering of Gibson, Glasbey & Elston (1994): at step j
    the remaining variable with the smallest conditional truncated mass comes next.
    Returns (L lower, l, u permuted, perm)."""
    Sig = np.array(Sig, float); l = np.array(l, float); u = np.array(u, float)
    d = l.size
    perm = np.arange(d)
    L = np.zeros((d, d)); z = np.zeros(d)
    for j in range(d):
        I = np.arange(j, d)
        s = np.diag(Sig)[I] - np.sum(L[I, :j] ** 2, axis=1)
        s = np.sqrt(np.where(s < 0, EPS, s))
        cols = L[I, :j] @ z[:j]
        pr = lnNpr((l[I] - cols) / s, (u[I] - cols) / s)
        k = j + int(np.argmin(pr))
        if k != j:                               # swap variables j <-> k everywhere
            Sig[[j, k], :] = Sig[[k, j], :]; Sig[:, [j, k]] = Sig[:, [k, j]]
            L[[j, k], :] = L[[k, j], :]
            l[[j, k]] = l[[k, j]]; u[[j, k]] = u[[k, j]]; perm[[j, k]] = perm[[k, j]]
        s = Sig[j, j] - np.sum(L[j, :j] ** 2)
        if s < -0.001:
            raise ValueError("Sigma is not positive semi-definite")
        if s < 0:
            s = EPS
        L[j, j] = np.sqrt(s)
        L[j + 1:, j] = (Sig[j + 1:, j] - L[j + 1:, :j] @ L[j, :j]) / L[j, j]
        tl = (l[j] - L[j, :j] @ z[:j]) / L[j, j]
        tu = (u[j] - L[j, :j] @ z[:j]) / L[j, j]
        w = lnNpr(tl, tu)
        z[j] = (np.exp(-0.5 * tl ** 2 - w) - np.exp(-0.5 * tu ** 2 - w)) / np.sqrt(2 * np.pi)
    return L, l, u, perm


def gradpsi(y, L, l, u):
    """Gradient and Jacobian of psi(x, mu) (L scaled, zero diagonal)."""
    d = u.size
    x = np.zeros(d); mu = np.zeros(d)
    x[:d - 1] = y[:d - 1]; mu[:d - 1] = y[d - 1:]
    c = np.zeros(d); c[1:] = L[1:, :] @ x
    lt = l - mu - c; ut = u - mu - c
    w = lnNpr(lt, ut)
    pl = np.exp(-0.5 * lt ** 2 - w) / np.sqrt(2 * np.pi)
    pu = np.exp(-0.5 * ut ** 2 - w) / np.sqrt(2 * np.pi)
    P = pl - pu
    dfdx = -mu[:d - 1] + (P @ L[:, :d - 1])
    dfdm = mu - x + P
    grad = np.concatenate([dfdx, dfdm[:d - 1]])
    lt = np.where(np.isinf(lt), 0.0, lt); ut = np.where(np.isinf(ut), 0.0, ut)
    dP = -P ** 2 + lt * pl - ut * pu
    DL = dP[:, None] * L
    mx = (-np.eye(d) + DL)[:d - 1, :d - 1]
    xx = (L.T @ DL)[:d - 1, :d - 1]
    Jac = np.block([[xx, mx.T], [mx, np.diag(1 + dP[:d - 1])]])
    return grad, Jac


def nleq(l, u, L):
    """Newton iteration for grad psi = 0 from the origin."""
    d = l.size
    y = np.zeros(2 * d - 2)
    for it in range(101):
        grad, Jac = gradpsi(y, L, l, u)
        if grad.size == 0:
            break
        y = y + np.linalg.solve(Jac, -grad)
        if np.sum(grad ** 2) <= 1e-10:
            break
    else:
        raise ValueError("Covariance matrix is ill-conditioned and method failed")
    return y


def psy(x, L, l, u, mu):
    """psi(x, mu) (L scaled, zero diagonal; x and mu padded with a final zero)."""
    d = u.size
    x = np.append(x[:d - 1], 0.0); mu = np.append(mu[:d - 1], 0.0)
    c = L @ x
    return float(np.sum(lnNpr(l - mu - c, u - mu - c) + 0.5 * mu ** 2 - x * mu))


def setup(l, u, Sig):
    """Everything mvrandn does before its accept-reject loop.  Returns a dict with
    L0 (scaled factor, zero diagonal), l, u (scaled, permuted), mu (tilt, last entry 0),
    psistar, Lfull (unscaled factor), order (inverse permutation)."""
    l = np.asarray(l, float).ravel(); u = np.asarray(u, float).ravel()
    if np.any(l >= u):
        raise ValueError("l, u, and Sig have to match in dimension with u>l")
    Lfull, lp, up, perm = cholperm(Sig, l, u)
    D = np.diag(Lfull).copy()
    Ls = Lfull / D[:, None]
    lp = lp / D; up = up / D
    L0 = Ls - np.eye(l.size)
    y = nleq(lp, up, L0)
    d = l.size
    x = y[:d - 1]; mu = np.append(y[d - 1:], 0.0)
    psistar = psy(x, L0, lp, up, mu)
    order = np.argsort(perm)
    return dict(L0=L0, l=lp, u=up, mu=mu, psistar=psistar, Lfull=Lfull, order=order, perm=perm, x=x)


def setup_device(l, u, Sig):
    """expt.setup on the GPU (lqg_expt_setup: permuted Cholesky in one CTA, Newton tilt through the SPD Schur
    complement on the FP64 tensor cores).  Same dict as setup()."""
    from . import _lib as L
    l = np.ascontiguousarray(np.asarray(l, float).ravel()); u = np.ascontiguousarray(np.asarray(u, float).ravel())
    d = l.size
    if np.any(l >= u):
        raise ValueError("l, u, and Sig have to match in dimension with u>l")
    Sf = L.f64(np.asarray(Sig, float))
    L0 = np.empty((d, d), order="F"); Lfull = np.empty((d, d), order="F")
    lo = np.empty(d); uo = np.empty(d); mu = np.empty(d); x = np.empty(d)
    perm = np.empty(d, dtype=np.int32)
    ps = L.C.c_double(0.0); it = L.C.c_int32(0)
    rc = L.lib().lqg_expt_setup(d, L.ptr(Sf), L.ptr(l), L.ptr(u), L.make_opts(mem=L.MEM_HOST), L.ptr(L0), L.ptr(lo), L.ptr(uo),
                                L.ptr(mu), L.ptr(x), L.C.byref(ps), perm.ctypes.data, L.ptr(Lfull), L.C.byref(it))
    if rc == 8:
        raise ValueError("Sigma is not positive semi-definite")
    if rc == 1:
        raise ValueError("Covariance matrix is ill-conditioned and method failed")
    L.check(rc, "lqg_expt_setup")
    perm = perm.astype(np.int64)
    return dict(L0=L0, l=lo, u=uo, mu=mu, psistar=float(ps.value), Lfull=Lfull, order=np.argsort(perm), perm=perm,
                x=x[:d - 1], newton_iterations=int(it.value))
