"""Host-side mirror of the reference's model layer: the callers of the hot path.

These are the R functions that BUILD the inputs of `tmvrnorm` (kernels, hat basis, inequality
systems, GP posterior, MAP).  In the reference they are interpreted R + LAPACK on the host and
they stay host code here (numpy): SURVEY.md §8 marks them "caller/setup", not kernels.  They
exist so that `simulate()` is usable end to end and so that benchmarks/tests can synthesise the
named configurations.  Names, arguments and error behaviour follow the R functions cited.

  R/... = /root/reference/dev/lineqGPR/R/...
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------- kernels.R


def k1gaussian(x1, x2, par):
    """R/kernels.R:61-85."""
    sigma2, theta = par[0], par[1]
    dist = np.subtract.outer(np.ravel(x1) / theta, np.ravel(x2) / theta)
    return sigma2 * np.exp(-0.5 * dist ** 2)


def k1matern32(x1, x2, par):
    """R/kernels.R (k1matern32)."""
    sigma2, theta = par[0], par[1]
    a = np.sqrt(3.0) * np.abs(np.subtract.outer(np.ravel(x1) / theta, np.ravel(x2) / theta))
    return sigma2 * (1 + a) * np.exp(-a)


def k1matern52(x1, x2, par):
    """R/kernels.R:195-224."""
    sigma2, theta = par[0], par[1]
    dist = np.subtract.outer(np.ravel(x1) / theta, np.ravel(x2) / theta)
    a = np.sqrt(5.0) * np.abs(dist)
    return sigma2 * (1 + a + (5.0 / 3.0) * dist ** 2) * np.exp(-a)


def k1exponential(x1, x2, par):
    """R/kernels.R:246-261."""
    sigma2, theta = par[0], par[1]
    return sigma2 * np.exp(-np.abs(np.subtract.outer(np.ravel(x1) / theta, np.ravel(x2) / theta)))


_KERNELS = {"gaussian": k1gaussian, "matern32": k1matern32, "matern52": k1matern52,
            "exponential": k1exponential}


def kernCompute(x1, x2=None, type="gaussian", par=(1.0, 0.2)):
    """R/kernels.R kernCompute dispatcher (1-D kernels only: all this path needs)."""
    if x2 is None:
        x2 = x1
    if type not in _KERNELS:
        raise ValueError(f'kernel "{type}" is not supported')
    return _KERNELS[type](x1, x2, np.asarray(par, float))


# --------------------------------------------------------------------------- hat basis


def basisCompute_lineqGP(x, u, d=1):
    """R/lineqGP.R:29-76.  Hat (piecewise-linear) basis; rows = points, columns = knots."""
    if d == 1:
        x = np.asarray(x, float).ravel()
        u = np.asarray(u[0] if isinstance(u, (list, tuple)) else u, float).ravel()
        m = u.size
        diff = x[:, None] - u[None, :]
        du_r = np.full(m, np.nan); du_r[:-1] = u[1:] - u[:-1]          # u[j+1]-u[j]  (:50)
        du_l = np.full(m, np.nan); du_l[1:] = u[1:] - u[:-1]           # u[j]-u[j-1]  (:52)
        right = diff > 1e-12                                            # :49
        left = diff < 0                                                 # :51
        with np.errstate(invalid="ignore"):
            dist = np.where(right, np.abs(diff / du_r[None, :]), np.where(left, np.abs(diff / du_l[None, :]), 0.0))
        if np.isnan(dist).any():                                        # Q9: R indexes u out of range -> NA -> error
            raise ValueError("basisCompute.lineqGP: x outside the knot range [u_1, u_m]")
        return np.where(dist <= 1.0, 1.0 - dist, 0.0)                   # :58-59
    x = np.asarray(x, float).reshape(-1, d)
    plist = [basisCompute_lineqGP(x[:, k], u[k], 1) for k in range(d)]  # :62-65
    Phi = plist[0]
    for k in range(1, d):                                               # :67-72 row-wise Kronecker
        Phi = (Phi[:, :, None] * plist[k][:, None, :]).reshape(x.shape[0], -1)
    return Phi


# --------------------------------------------------------------------------- inequality systems


def bounds2lineqSys(d=None, l=0.0, u=1.0, A=None, lineqSysType="twosides", rmInf=True):
    """R/lineqGPutils.R:353-390."""
    if A is None:
        A = np.eye(d)
    if d is None:
        d = A.shape[0]
    l = np.atleast_1d(np.asarray(l, float)).ravel()
    u = np.atleast_1d(np.asarray(u, float)).ravel()
    if l.size not in (1, d) or u.size not in (1, d):
        raise ValueError("The bounds have to be d-dimensionals")
    if A.shape[0] != d:
        raise ValueError("nrow(A) has to be equal to d")
    if np.any(l >= u):
        raise ValueError('The elements from "u" has to be greater than the elements from "l"')
    if l.size == 1:
        l = np.repeat(l, d)
    if u.size == 1:
        u = np.repeat(u, d)
    if lineqSysType == "twosides":
        if rmInf:
            keep = ~((l == -np.inf) & (u == np.inf))
            A, l, u = A[keep], l[keep], u[keep]
        return dict(A=A, l=l, u=u)
    if lineqSysType == "oneside":
        g = np.concatenate([-l, u])
        M = np.vstack([A, -A])
        if rmInf:
            keep = ~(g == np.inf)
            g, M = g[keep], M[keep]
        return dict(M=M, g=g)
    raise ValueError(f'type of linear system "{lineqSysType}" is not supported')


def _base_system_1d(m, constrType, l, u):
    """The d == 1 branch of lineqGPSys (R/lineqGP.R:119-163): returns (A, l, u) before bounds2lineqSys."""
    l = np.atleast_1d(np.asarray(l, float)); u = np.atleast_1d(np.asarray(u, float))
    if constrType in ("boundedness", "none"):
        return np.eye(m), l, u
    if constrType == "monotonicity":
        if l.size == 1: l = np.concatenate([[-np.inf], np.repeat(l, m - 1)])
        if u.size == 1: u = np.concatenate([[np.inf], np.repeat(u, m - 1)])
        A = np.eye(m)
        A[np.arange(1, m), np.arange(0, m - 1)] = -1.0
        return A, l, u
    if constrType == "decreasing":
        if l.size == 1: l = np.concatenate([[-np.inf], np.repeat(l, m - 1)])
        if u.size == 1: u = np.concatenate([[np.inf], np.repeat(u, m - 1)])
        A = -np.eye(m)
        A[np.arange(1, m), np.arange(0, m - 1)] = 1.0
        A[0, 0] = 1.0
        return A, l, u
    if constrType == "convexity":
        if l.size == 1: l = np.concatenate([[-np.inf, -np.inf], np.repeat(l, m - 2)])
        if u.size == 1: u = np.concatenate([[np.inf, np.inf], np.repeat(u, m - 2)])
        A = np.eye(m)
        A[np.arange(2, m), np.arange(0, m - 2)] = 1.0
        A[np.arange(2, m), np.arange(1, m - 1)] = -2.0
        return A, l, u
    raise ValueError(f"'arg' should be one of the supported constraint types, got {constrType!r}")


def lineqGPSys(m, constrType="boundedness", l=-np.inf, u=np.inf, A=None, d=None,
               lineqSysType="twosides", constrIdx=None, rmInf=True):
    """R/lineqGP.R:111-196."""
    m_arr = np.atleast_1d(np.asarray(m, int))
    if d is None:
        d = m_arr.size
    if constrType == "linear":
        return bounds2lineqSys(A.shape[0], l, u, A, lineqSysType, rmInf)
    if d == 1:
        A1, l1, u1 = _base_system_1d(int(m_arr[0]), constrType, l, u)
        return bounds2lineqSys(A1.shape[0], l1, u1, A1, lineqSysType, rmInf)
    if constrIdx is None:
        constrIdx = list(range(d))
    base = []
    for mk in m_arr:                                                    # :166-170 (twosides, rmInf=FALSE)
        Ab, lb_, ub_ = _base_system_1d(int(mk), constrType, l, u)
        s = bounds2lineqSys(Ab.shape[0], lb_, ub_, Ab, "twosides", False)
        base.append((s["A"], s["l"], s["u"]))
    Arows, ls, us = [], [], []
    for k in constrIdx:                                                 # :182-189 Kronecker structure
        Ak = np.ones((1, 1)); lk = np.ones(1); uk = np.ones(1)
        for j in range(d):
            if j == k:
                Ak = np.kron(Ak, base[j][0])
                with np.errstate(invalid="ignore"):
                    lk = np.kron(lk, base[j][1]); uk = np.kron(uk, base[j][2])
            else:
                Ak = np.kron(Ak, np.eye(int(m_arr[j])))
                lk = np.kron(lk, np.ones(int(m_arr[j]))); uk = np.kron(uk, np.ones(int(m_arr[j])))
        Arows.append(Ak); ls.append(lk); us.append(uk)
    A_ = np.vstack(Arows); l_ = np.concatenate(ls); u_ = np.concatenate(us)
    return bounds2lineqSys(A_.shape[0], l_, u_, A_, lineqSysType, rmInf)


# --------------------------------------------------------------------------- QP (solve.QP stand-in)


def solve_qp(D, dvec, A, b, tol=1e-10, max_iter=200, polish=True):
    """min 1/2 x'Dx - dvec'x  s.t.  A x >= b   (quadprog::solve.QP(Dmat, dvec, t(A), b) as used at
    R/lineqGP.R:530, R/lineqAGP.R:431).  quadprog is not available here; this is a Mehrotra
    predictor-corrector interior-point method (dense, O(m^3) per iteration).
    Merit phi = max(|rd|/(1e3 scale_d), |rp|/scale_p, mu/scale_p); converged when phi < tol.
    With the normal-equations form the attainable dual residual grows like eps |A|^2 lam^2 / mu,
    so close to the solution phi can bottom out just above tol and then rise again: the best
    iterate is kept, and when phi has not improved for 5 iterations it is returned provided
    phi_best < 1e3 tol (the rounding floor of this problem was reached).
    polish: finish with an active-set solve (polish_qp) so that, like quadprog, the exact
    minimiser is returned."""
    D = np.asarray(D, float); dvec = np.asarray(dvec, float).ravel()
    A = np.asarray(A, float); b = np.asarray(b, float).ravel()
    n, mc = D.shape[0], A.shape[0]
    x = np.linalg.solve(D, dvec)
    if mc == 0:
        return x
    s = np.maximum(A @ x - b, 1.0)
    lam = np.ones(mc)
    scale_d = max(1.0, np.abs(dvec).max())
    fin = np.isfinite(b)
    scale_p = max(1.0, np.abs(b[fin]).max() if fin.any() else 1.0)
    best_phi, best_x, best_sl, stall = np.inf, x, (s, lam), 0
    for _ in range(max_iter):
        rd = D @ x - dvec - A.T @ lam
        rp = A @ x - s - b
        mu = s @ lam / mc
        phi = max(np.abs(rd).max() / (1e3 * scale_d), np.abs(rp).max() / scale_p, mu / scale_p)
        if phi < tol:
            xp = polish_qp(D, dvec, A, b, s, lam) if polish else None
            return x if xp is None else xp
        if phi < best_phi:
            best_phi, best_x, best_sl, stall = phi, x.copy(), (s.copy(), lam.copy()), 0
        else:
            stall += 1
            if stall >= 5:
                break
        W = lam / s
        H = D + (A.T * W) @ A
        try:
            L = np.linalg.cholesky(H)
        except np.linalg.LinAlgError:
            L = np.linalg.cholesky(H + 1e-12 * np.trace(H) / n * np.eye(n))

        def step(rc):
            # solves the reduced KKT system for (dx, ds, dlam) given complementarity residual rc
            rhs = -rd + A.T @ ((-rc - lam * rp) / s)
            dx = np.linalg.solve(L.T, np.linalg.solve(L, rhs))
            ds = A @ dx + rp
            dl = (-rc - lam * ds) / s
            return dx, ds, dl

        dx, ds, dl = step(s * lam)
        def amax(v, dv):
            neg = dv < 0
            return min(1.0, (-v[neg] / dv[neg]).min()) if neg.any() else 1.0
        a_aff = min(amax(s, ds), amax(lam, dl))
        mu_aff = (s + a_aff * ds) @ (lam + a_aff * dl) / mc
        sigma = (mu_aff / mu) ** 3
        dx, ds, dl = step(s * lam + ds * dl - sigma * mu)
        a = 0.995 * min(amax(s, ds), amax(lam, dl))
        x = x + a * dx; s = s + a * ds; lam = lam + a * dl
    if best_phi < 1e3 * tol:
        xp = polish_qp(D, dvec, A, b, *best_sl) if polish else None
        return best_x if xp is None else xp
    return x


def polish_qp(D, dvec, A, b, s, lam, max_rounds=8):
    """Active-set polish of an interior-point iterate: quadprog::solve.QP is an active-set method
    and returns the exact minimiser, an interior-point iterate is only O(sqrt(mu)) close on weakly
    active rows.  Guess the active set (lam_i > s_i), solve the equality-constrained problem
    through the Schur complement  (A_a D^-1 A_a') lam_a = b_a - A_a D^-1 dvec,
    x = D^-1 (dvec + A_a' lam_a),  and check the KKT conditions of the full problem; rows that
    are violated are added, rows with a negative multiplier dropped, a few times.  Returns the
    exact minimiser, or None when no consistent active set was found (the caller keeps the
    interior-point iterate)."""
    D = np.asarray(D, float); A = np.asarray(A, float); b = np.asarray(b, float).ravel()
    dvec = np.asarray(dvec, float).ravel()
    Lc = np.linalg.cholesky(D)
    solveD = lambda R: np.linalg.solve(Lc.T, np.linalg.solve(Lc, R))
    x0 = solveD(dvec)
    scale_p = max(1.0, np.abs(b).max(initial=0.0))
    act = lam > s
    for _ in range(max_rounds):
        idx = np.flatnonzero(act)
        if idx.size == 0:
            xp, lam_a = x0, np.zeros(0)
        else:
            Aa = A[idx]
            Y = solveD(Aa.T)                                   # n x k
            S = Aa @ Y
            try:
                Ls = np.linalg.cholesky((S + S.T) / 2)
            except np.linalg.LinAlgError:
                return None                                    # dependent active rows
            lam_a = np.linalg.solve(Ls.T, np.linalg.solve(Ls, b[idx] - Aa @ x0))
            xp = x0 + Y @ lam_a
        r = A @ xp - b
        viol = r < -1e-12 * scale_p
        neg = lam_a < -1e-10 * max(1.0, np.abs(lam_a).max(initial=0.0))
        if not viol.any() and not neg.any():
            return xp
        new = act.copy()
        new[idx[neg]] = False
        new[viol] = True
        if np.array_equal(new, act):
            return None
        act = new
    return None


def _chol2inv(S):
    L = np.linalg.cholesky(S)
    Li = np.linalg.inv(L)
    return Li.T @ Li


class _HostAlgebra:
    """The dense set-up algebra of predict.* as R runs it (LAPACK on the host)."""
    chol2inv = staticmethod(_chol2inv)
    solve_qp = staticmethod(solve_qp)

    @staticmethod
    def min_eig_nonpositive(S):
        return np.linalg.eigvalsh((S + S.T) / 2).min() <= 0


class _DeviceAlgebra:
    """The same three operations through the C-ABI set-up stage (csrc/lqg_linalg.cu, lqg_qp.cu);
    raises when the CUDA library or a device is missing (no fallback)."""

    @staticmethod
    def chol2inv(S):
        from . import setup_stage as st
        S = np.asarray(S, float)
        inv = st.solve_spd(S, np.eye(S.shape[0]))
        return (inv + inv.T) / 2

    @staticmethod
    def min_eig_nonpositive(S):
        from . import setup_stage as st
        return not st.is_positive_definite(S)

    @staticmethod
    def solve_qp(D, dvec, A, b):
        from . import setup_stage as st
        return st.solve_qp_device(D, dvec, A, b)


def _algebra(setup):
    if setup == "host":
        return _HostAlgebra
    if setup == "device":
        return _DeviceAlgebra
    raise ValueError(f"unknown setup {setup!r} (host | device)")


# --------------------------------------------------------------------------- lineqGP


def create_lineqGP(x, y, constrType, m=None):
    """create.lineqGP, R/lineqGP.R:237-293."""
    x = np.asarray(x, float)
    if x.ndim == 1:
        x = x[:, None]
    y = np.asarray(y, float).reshape(-1, 1)
    d = x.shape[1]
    constrType = [constrType] if isinstance(constrType, str) else list(constrType)
    if m is None:
        m = [int(10 * y.size / d)] * d
    elif np.size(m) == 1:
        m = [int(np.ravel(m)[0])] * d
    elif np.size(m) != d:
        raise ValueError("The length of 'm' has to be equal to the input dimension")
    m = [int(v) for v in np.ravel(m)]
    u = [np.linspace(0.0, 1.0, mk) for mk in m]
    kernParam = dict(par=np.array([1.0] + [0.2] * d), type="gaussian",
                     nugget=1e-7 * float(np.std(y, ddof=1)))                # :261-262
    localParam = dict(m=m, sampler="ExpT", samplingParams=dict(thinning=1, **{"burn.in": 1}, scale=0.1))
    model = dict(cls="lineqGP", x=x, y=y, constrType=constrType, ulist=u, d=d,
                 constrIdx=list(range(d)), varnoise=0.0, localParam=localParam, kernParam=kernParam)
    bounds = []
    if "none" in constrType:
        bounds = [[-np.inf, np.inf]]
    else:
        if "boundedness" in constrType:                                      # :275-278 (Q10 typo kept)
            bounds.append([y.min() - 0.05 * abs(y.max() - y.min()), y.max() + 0.05 * abs(y.max() - y.max())])
        for ct in ("monotonicity", "decreasing", "convexity"):
            if ct in constrType:
                bounds.append([0.0, np.inf])
    model["bounds"] = np.array(bounds, float)
    if "linear" in constrType:
        mt = int(np.prod(m))
        model["Lambda"] = np.eye(mt); model["lb"] = np.zeros(mt); model["ub"] = np.full(mt, np.inf)
    return model


def augment_lineqGP(model):
    """augment.lineqGP, R/lineqGP.R:342-434."""
    model = dict(model)
    kp = dict(model["kernParam"]); kp.setdefault("nugget", 0.0); model["kernParam"] = kp
    bounds = np.asarray(model.get("bounds", [[0.0, np.inf]]), float).reshape(-1, 2)
    x = model["x"]
    d = model["d"]
    u = model["ulist"]
    if d == 1:
        u1 = np.asarray(u[0] if isinstance(u, (list, tuple)) else u, float).ravel()
        m = [u1.size]
        Gamma = kernCompute(u1, u1, kp["type"], kp["par"])
        ubasis = u1
    else:
        m = [len(uk) for uk in u]
        par = np.asarray(kp["par"], float)
        if par.size == 2:
            par = np.concatenate([[par[0]], np.repeat(par[1], d)])
        Gamma = np.ones((1, 1))
        for j in range(d):                                               # :387-388 Kronecker prior
            Gamma = np.kron(Gamma, kernCompute(u[j], u[j], kp["type"], [par[0], par[j + 1]]))
        ubasis = u
    model["localParam"] = dict(model["localParam"], m=m)
    model["Gamma"] = Gamma
    model["Phi"] = basisCompute_lineqGP(x, ubasis, x.shape[1])               # :391
    Ms, gs, Ls, lbs, ubs, mvec = [], [], [], [], [], []
    for i, ct in enumerate(model["constrType"]):                             # :398-424
        if ct == "linear":
            if "Lambda" not in model:
                raise ValueError("matrix Lambda is not defined")
            lsys = lineqGPSys(model["Lambda"].shape[0], ct, model["lb"], model["ub"], model["Lambda"], lineqSysType="oneside")
            lsys2 = lineqGPSys(model["Lambda"].shape[0], ct, model["lb"], model["ub"], model["Lambda"], rmInf=False)
        else:
            lsys = lineqGPSys(m, ct, bounds[i, 0], bounds[i, 1], d=d, constrIdx=model["constrIdx"], lineqSysType="oneside")
            lsys2 = lineqGPSys(m, ct, bounds[i, 0], bounds[i, 1], d=d, constrIdx=model["constrIdx"], rmInf=False)
        Ms.append(lsys["M"]); gs.append(-lsys["g"])
        Ls.append(lsys2["A"]); lbs.append(lsys2["l"]); ubs.append(lsys2["u"])
        mvec.append(lsys2["A"].shape[0])
    model["Lambda"] = np.vstack(Ls); model["lb"] = np.concatenate(lbs); model["ub"] = np.concatenate(ubs)
    model["lineqSys"] = dict(M=np.vstack(Ms), g=np.concatenate(gs))
    model["localParam"]["mvec"] = mvec
    model["localParam"]["u"] = ubasis
    return model


def predict_lineqGP(model, xtest, solve_map=True, setup="host"):
    """predict.lineqGP, R/lineqGP.R:493-535.  solve_map=False skips the QP (xi.map = None).
    setup="device" runs chol2inv, the positive-definiteness test and the QP on the GPU."""
    la = _algebra(setup)
    model = augment_lineqGP(model)
    xtest = np.asarray(xtest, float)
    pred = dict(Lambda=model["Lambda"], lb=model["lb"], ub=model["ub"])
    pred["Phi.test"] = basisCompute_lineqGP(xtest, model["localParam"]["u"], model["d"])
    Phi, Gamma, y = model["Phi"], model["Gamma"], model["y"]
    GammaPhit = Gamma @ Phi.T                                                # :514
    PhiGammaPhit = Phi @ GammaPhit
    if model["varnoise"] != 0:
        PhiGammaPhit = PhiGammaPhit + model["varnoise"] * np.eye(Phi.shape[0])
    inv = la.chol2inv(PhiGammaPhit)                                          # :518
    pred["mu"] = (GammaPhit @ inv @ y).ravel()
    Sigma = Gamma - GammaPhit @ inv @ GammaPhit.T                            # :520
    if la.min_eig_nonpositive(Sigma):                                        # :527-528
        Sigma = Sigma + model["kernParam"]["nugget"] * np.eye(Sigma.shape[0])
    pred["Sigma"] = Sigma
    if solve_map:
        invSigma = la.chol2inv(Sigma)
        pred["xi.map"] = la.solve_qp(invSigma, pred["mu"] @ invSigma, model["lineqSys"]["M"], model["lineqSys"]["g"])
    else:
        pred["xi.map"] = None
    pred["model"] = model
    return pred


# --------------------------------------------------------------------------- lineqAGP


def create_lineqAGP(x, y, constrType, m=None):
    """create.lineqAGP, R/lineqAGP.R:46-112."""
    x = np.asarray(x, float)
    if x.ndim == 1:
        x = x[:, None]
    y = np.asarray(y, float).reshape(-1, 1)
    d = x.shape[1]
    constrType = [constrType] * d if isinstance(constrType, str) else list(constrType)
    if m is None:
        m = [int(10 * y.size / d)] * d
    elif np.size(m) == 1:
        m = [int(np.ravel(m)[0])] * d
    elif np.size(m) == d:
        m = [int(v) for v in np.ravel(m)]     # MaxMod models carry per-dimension knot counts
    else:
        raise ValueError("The length of 'm' has to be equal to the input dimension")
    u = [np.linspace(0.0, 1.0, mk) for mk in m]
    kernParam = [dict(par=np.array([1.0, 0.1]), type="matern52") for _ in range(d)]   # :82
    constrParam = []
    for k in range(d):
        ct = constrType[k]
        if ct == "boundedness":
            constrParam.append(dict(bounds=np.array([y.min() - 0.05 * abs(y.max() - y.min()), y.max()])))
        elif ct in ("monotonicity", "decreasing", "convexity"):
            constrParam.append(dict(bounds=np.array([0.0, np.inf])))
        elif ct == "none":
            constrParam.append(dict(bounds=np.array([-np.inf, np.inf])))
        else:
            raise ValueError(f"constraint type {ct!r} is not supported for lineqAGP here")
    return dict(cls="lineqAGP", x=x, y=y, constrType=constrType, ulist=u, d=d, nugget=1e-9,
                constrIdx=[k for k in range(d) if constrType[k] != "none"], constrParam=constrParam,
                varnoise=0.0, kernParam=kernParam,
                localParam=dict(m=m, sampler="ExpT", ngroups=d,
                                samplingParams=dict(thinning=1, **{"burn.in": 1}, scale=0.1)))


def _bdiag(mats):
    r = sum(a.shape[0] for a in mats); c = sum(a.shape[1] for a in mats)
    out = np.zeros((r, c)); i = j = 0
    for a in mats:
        out[i:i + a.shape[0], j:j + a.shape[1]] = a
        i += a.shape[0]; j += a.shape[1]
    return out


def augment_lineqAGP(model):
    """augment.lineqAGP, R/lineqAGP.R:167-253."""
    model = dict(model)
    model.setdefault("nugget", 0.0)
    x, u = model["x"], model["ulist"]
    D = model["localParam"]["ngroups"]
    m = [len(uk) for uk in u]
    model["localParam"] = dict(model["localParam"], m=m, mtotal=int(sum(m)))
    model["Gamma"] = [kernCompute(u[k], u[k], model["kernParam"][k]["type"], model["kernParam"][k]["par"]) for k in range(D)]
    model["Phi"] = [basisCompute_lineqGP(x[:, k], u[k]) for k in range(D)]
    Ms, gs, Ls, lbs, ubs = [], [], [], [], []
    for k in range(D):                                                       # :195-222
        b = model["constrParam"][k]["bounds"]
        lsys = lineqGPSys(m[k], model["constrType"][k], b[0], b[1], lineqSysType="oneside")
        lsys2 = lineqGPSys(m[k], model["constrType"][k], b[0], b[1], rmInf=False)
        Ms.append(lsys["M"]); gs.append(-lsys["g"])
        Ls.append(lsys2["A"]); lbs.append(lsys2["l"]); ubs.append(lsys2["u"])
    model["lineqSys"] = dict(MAll=_bdiag(Ms), gAll=np.concatenate(gs), LambdaAll=_bdiag(Ls),
                             lb=np.concatenate(lbs), ub=np.concatenate(ubs))   # :229-249
    return model


def predict_lineqAGP(model, xtest, solve_map=True, setup="host"):
    """predict.lineqAGP, R/lineqAGP.R:329-443 (the mt >= nt branch and the mt < nt Woodbury branch).
    setup="device": chol2inv and the QP on the GPU."""
    la = _algebra(setup)
    model = augment_lineqAGP(model)
    xtest = np.asarray(xtest, float).reshape(-1, model["d"])
    D = model["localParam"]["ngroups"]
    pred = dict()
    pred["Phi.test"] = [basisCompute_lineqGP(xtest[:, k], model["ulist"][k]) for k in range(D)]
    pred["PhiAll.test"] = np.hstack(pred["Phi.test"])                        # :343-348
    bigPhi = np.hstack(model["Phi"])
    bigGamma = _bdiag(model["Gamma"])
    y = model["y"]
    nt, mt = y.size, model["localParam"]["mtotal"]
    if mt < nt and model["varnoise"] > 0:                                    # :365-376
        Lg = _bdiag([np.linalg.cholesky(G + 0.0 * np.eye(G.shape[0])) for G in model["Gamma"]])
        PL = bigPhi @ Lg
        I = model["varnoise"] * np.eye(mt) + PL.T @ PL
        Lc = np.linalg.cholesky(I)
        Ls = np.linalg.solve(Lc, PL.T)
        inv = (np.eye(nt) - Ls.T @ Ls) / model["varnoise"]
    else:                                                                    # :377-381
        PGP = bigPhi @ bigGamma @ bigPhi.T + model["nugget"] * np.eye(nt)
        inv = la.chol2inv(PGP + model["varnoise"] * np.eye(nt))
    temp = bigGamma @ bigPhi.T @ inv                                         # :403
    pred["muAll"] = (temp @ y).ravel()
    C = temp @ bigPhi @ bigGamma
    C = (C + C.T) / 2
    pred["SigmaAll"] = bigGamma - C                                          # :407
    if solve_map:
        invS = la.chol2inv(pred["SigmaAll"] + model["nugget"] * np.eye(mt))  # :429
        pred["xiAll.map"] = la.solve_qp(invS, pred["muAll"] @ invS, model["lineqSys"]["MAll"], model["lineqSys"]["gAll"])
    else:
        pred["xiAll.map"] = None
    pred["model"] = model
    return pred
