"""ORACLE — TEST INFRASTRUCTURE ONLY.

Independent numpy restatement of the minimax-tilting sampler TruncatedNormal::mvrandn, the body
of tmvrnorm.ExpT (call site R/lineqGPsamplers.R:183).  TruncatedNormal (CRAN, imported unversioned,
lineqGPR DESCRIPTION:12) is NOT vendored under /root/reference: this follows the published
algorithm — Botev (2017) JRSS-B 79(1):125-148 — so PARITY WITH R IS UNPINNED; the oracle pins the
GPU kernel to this restatement draw for draw (same Philox-addressed uniforms) and to analytic
anchors (independent case: psi* = log P exactly; 1-D: scipy truncnorm).

Written separately from lineqgpr_b200/expt.py (scalar loops, scipy.stats), on purpose."""
from __future__ import annotations

import numpy as np
from scipy.stats import norm

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
TAG_EXPT = 0x45585054
MASK = 0xFFFFFFFF


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox-4x32-10 (Salmon et al. 2011), scalar Python ints."""
    for _ in range(10):
        p0 = M0 * c0; p1 = M1 * c2
        hi0, lo0 = (p0 >> 32) & MASK, p0 & MASK
        hi1, lo1 = (p1 >> 32) & MASK, p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK, lo1, (hi0 ^ c3 ^ k1) & MASK, lo0
        k0 = (k0 + W0) & MASK; k1 = (k1 + W1) & MASK
    return c0, c1, c2, c3


def u01(hi, lo):
    return ((((hi << 32) | lo) >> 11) + 0.5) / 9007199254740992.0


class Draws:
    """Uniform stream of one (proposal, coordinate): pair j = Philox(ctr=(j, p_lo, p_hi, k))."""

    def __init__(self, seed, p, k):
        self.seed, self.p, self.k, self.j = seed, p, k, 0

    def next2(self):
        r = philox4x32_10(self.j, self.p & MASK, (self.p >> 32) & MASK, self.k,
                          self.seed & MASK, ((self.seed >> 32) & MASK) ^ TAG_EXPT)
        self.j += 1
        return u01(r[0], r[1]), u01(r[2], r[3])


def lnNpr(a, b):
    if a > 0:
        pa, pb = norm.logsf(a), norm.logsf(b)
        return pa + np.log1p(-np.exp(pb - pa))
    if b < 0:
        pa, pb = norm.logcdf(a), norm.logcdf(b)
        return pb + np.log1p(-np.exp(pa - pb))
    return np.log1p(-norm.cdf(a) - norm.sf(b))


def ntail(l, u, g):
    c = l * l / 2; f = np.expm1(c - u * u / 2) if np.isfinite(u) else -1.0
    while True:
        u1, u2 = g.next2()
        x = c - np.log(1 + u1 * f)
        if u2 * u2 * x <= c:
            return np.sqrt(2 * x)


def tn(l, u, g):
    if abs(u - l) > 2.05:
        while True:
            u1, u2 = g.next2()
            x = np.sqrt(-2 * np.log(u1)) * np.cos(2 * np.pi * u2)
            if l < x < u:
                return x
    u1, _ = g.next2()
    pl, pu = norm.sf(l), norm.sf(u)
    return norm.isf(pl - (pl - pu) * u1)


def trandn(l, u, g):
    a = 0.66
    if l > a:
        return ntail(l, u, g)
    if u < -a:
        return -ntail(-u, -l, g)
    return tn(l, u, g)


def cholperm(Sig, l, u):
    Sig = np.array(Sig, float); l = np.array(l, float); u = np.array(u, float)
    d = len(l); perm = list(range(d)); L = np.zeros((d, d)); z = np.zeros(d)
    eps = np.finfo(float).eps
    for j in range(d):
        best, bk = np.inf, j
        for i in range(j, d):
            s = Sig[i, i] - sum(L[i, t] ** 2 for t in range(j))
            s = np.sqrt(eps if s < 0 else s)
            c = sum(L[i, t] * z[t] for t in range(j))
            pr = lnNpr((l[i] - c) / s, (u[i] - c) / s)
            if pr < best:
                best, bk = pr, i
        k = bk
        if k != j:
            Sig[[j, k], :] = Sig[[k, j], :]; Sig[:, [j, k]] = Sig[:, [k, j]]
            L[[j, k], :] = L[[k, j], :]
            l[[j, k]] = l[[k, j]]; u[[j, k]] = u[[k, j]]
            perm[j], perm[k] = perm[k], perm[j]
        s = Sig[j, j] - sum(L[j, t] ** 2 for t in range(j))
        if s < -0.001:
            raise ValueError("Sigma is not positive semi-definite")
        L[j, j] = np.sqrt(eps if s < 0 else s)
        for i in range(j + 1, d):
            L[i, j] = (Sig[i, j] - sum(L[i, t] * L[j, t] for t in range(j))) / L[j, j]
        c = sum(L[j, t] * z[t] for t in range(j))
        tl, tu = (l[j] - c) / L[j, j], (u[j] - c) / L[j, j]
        w = lnNpr(tl, tu)
        z[j] = (np.exp(-0.5 * tl ** 2 - w) - np.exp(-0.5 * tu ** 2 - w)) / np.sqrt(2 * np.pi)
    return L, l, u, np.array(perm)


def psi_and_grad(y, L, l, u):
    """psi(x, mu) and its gradient, by the definition (Botev 2017, eq. 9-11)."""
    d = len(u)
    x = np.append(y[:d - 1], 0.0); mu = np.append(y[d - 1:], 0.0)
    c = L @ x
    lt, ut = l - mu - c, u - mu - c
    w = np.array([lnNpr(a, b) for a, b in zip(lt, ut)])
    psi = float(np.sum(w + 0.5 * mu ** 2 - x * mu))
    pl = np.exp(-0.5 * lt ** 2 - w) / np.sqrt(2 * np.pi); pu = np.exp(-0.5 * ut ** 2 - w) / np.sqrt(2 * np.pi)
    P = pl - pu
    gx = -mu[:d - 1] + L[:, :d - 1].T @ P
    gm = (mu - x + P)[:d - 1]
    return psi, np.concatenate([gx, gm])


def solve_tilt(l, u, L):
    """Root of grad psi by scipy's hybrid Powell solver (the reference uses a plain Newton loop)."""
    from scipy.optimize import root
    d = len(l)
    if d == 1:
        return np.zeros(0)
    sol = root(lambda y: psi_and_grad(y, L, l, u)[1], np.zeros(2 * d - 2), method="hybr", tol=1e-13)
    if not sol.success and np.abs(sol.fun).max() > 1e-8:
        raise ValueError("tilting equations not solved")
    return sol.x


def setup(l, u, Sig):
    Lfull, lp, up, perm = cholperm(Sig, l, u)
    D = np.diag(Lfull).copy()
    L0 = Lfull / D[:, None] - np.eye(len(l))
    lp, up = lp / D, up / D
    y = solve_tilt(lp, up, L0)
    d = len(l)
    mu = np.append(y[d - 1:], 0.0)
    psistar = psi_and_grad(y, L0, lp, up)[0] if d > 1 else float(lnNpr(lp[0], up[0]))
    return dict(L0=L0, l=lp, u=up, mu=mu, psistar=psistar, Lfull=Lfull, order=np.argsort(perm), perm=perm)


def mvrandn_proposals(su, n_proposals, seed, first=0):
    """n tilted proposals with the kernel's Philox addressing.  Returns (Z d x n, logpr, accept)."""
    L0, l, u, mu, psistar = su["L0"], su["l"], su["u"], su["mu"], su["psistar"]
    d = len(l)
    Z = np.zeros((d, n_proposals)); logpr = np.zeros(n_proposals); acc = np.zeros(n_proposals, bool)
    for p in range(n_proposals):
        gp = first + p
        lp = 0.0
        for k in range(d):
            col = float(L0[k, :k] @ Z[:k, p])
            tl, tu = l[k] - mu[k] - col, u[k] - mu[k] - col
            z = mu[k] + trandn(tl, tu, Draws(seed, gp, k))
            Z[k, p] = z
            lp += lnNpr(tl, tu) + 0.5 * mu[k] ** 2 - mu[k] * z
        ua, _ = Draws(seed, gp, d).next2()
        logpr[p] = lp
        acc[p] = -np.log(ua) > psistar - lp
    return Z, logpr, acc


def mvrandn(l, u, Sig, n, seed, su=None, max_proposals=10 ** 6):
    """First n accepted proposals in proposal order, back-transformed: the mvrandn result."""
    su = su or setup(l, u, Sig)
    cols, used = [], 0
    while len(cols) < n and used < max_proposals:
        Z, _, acc = mvrandn_proposals(su, 256, seed, first=used)
        for p in np.nonzero(acc)[0]:
            if len(cols) < n:
                cols.append((used + p, Z[:, p]))
        used += 256
    idx = np.array([c[0] for c in cols]); Zs = np.stack([c[1] for c in cols], axis=1)
    x = (su["Lfull"] @ Zs)[su["order"], :]
    return x, Zs, idx
