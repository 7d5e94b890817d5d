"""Set-up stage on the device: the dense FP64 algebra the reference
runs on the host either side of the sampler call, through the C-ABI entry points of
csrc/lqg_linalg.cu.  Same names/meaning as the R expressions they replace; numpy in, numpy out.
No CPU fallback: every function raises when the CUDA library or a device is missing."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


def whiten_box_device(Sigma):
    """Ri of tmg::rtmg for M = solve(Sigma) (tmg:R/tmg.R:15-28), i.e. the upper-triangular U with
    U U' = Sigma, by a reversed blocked Cholesky factorisation on the GPU (lqg_whiten_box)."""
    S = L.f64(Sigma)
    d = S.shape[0]
    if S.shape != (d, d):
        raise ValueError("Sigma must be square")
    U = np.empty((d, d), order="F")
    rc = L.lib().lqg_whiten_box(d, L.ptr(S), L.make_opts(), L.ptr(U), None)
    if rc == 8:
        raise ValueError("Error: M must be positive definite.")          # tmg:R/tmg.R:17-20
    L.check(rc, "whiten_box")
    return U


def chol(A):
    """Lower Cholesky factor L (L L' = (A + A')/2); ValueError when A is not positive definite."""
    A = L.f64(A)
    d = A.shape[0]
    Lf = np.empty((d, d), order="F")
    bad = C.c_int32(0)
    rc = L.lib().lqg_chol(d, L.ptr(A), L.make_opts(), L.ptr(Lf), C.byref(bad))
    if rc == 8:
        raise ValueError(f"matrix is not positive definite (pivot {bad.value})")
    L.check(rc, "chol")
    return Lf


def is_positive_definite(A):
    """The test `min(eigen(A)$values) > 0` of tmg:R/tmg.R:17 and R/lineqGP.R:527-528, decided by
    whether the Cholesky factorisation runs through."""
    A = L.f64(A)
    bad = C.c_int32(0)
    rc = L.lib().lqg_chol(A.shape[0], L.ptr(A), L.make_opts(), None, C.byref(bad))
    if rc == 8:
        return False
    L.check(rc, "is_positive_definite")
    return True


def solve_spd(A, B):
    """solve(A, B) for a symmetric positive definite A."""
    A = L.f64(A)
    B2 = L.f64(np.asarray(B, float).reshape(A.shape[0], -1))
    X = np.empty(B2.shape, order="F")
    rc = L.lib().lqg_solve_spd(A.shape[0], B2.shape[1], L.ptr(A), L.ptr(B2), L.make_opts(), L.ptr(X))
    if rc == 8:
        raise ValueError("matrix is not positive definite")
    L.check(rc, "solve_spd")
    return X.reshape(np.shape(B))


def eta_params(Lambda, mu, xi_map, Sigma, nugget=0.0):
    """(Lambda mu, Lambda xi_map, Lambda Sigma Lambda' + nugget I): R/lineqGP.R:611-614."""
    Lam = L.f64(Lambda)
    d, m = Lam.shape
    S = L.f64(Sigma)
    mu = np.ascontiguousarray(mu, dtype=np.float64)
    xi_map = np.ascontiguousarray(xi_map, dtype=np.float64)
    eta_mu, eta_map = np.empty(d), np.empty(d)
    Se = np.empty((d, d), order="F")
    rc = L.lib().lqg_eta_params(d, m, L.ptr(Lam), L.ptr(mu), L.ptr(xi_map), L.ptr(S), float(nugget),
                                L.make_opts(), L.ptr(eta_mu), L.ptr(eta_map), L.ptr(Se))
    L.check(rc, "eta_params")
    return eta_mu, eta_map, Se


def lambda_pinv_device(Lambda):
    """The operator of qr.solve(Lambda, .) (R/lineqGP.R:639): m x d."""
    Lam = L.f64(Lambda)
    d, m = Lam.shape
    Lp = np.empty((m, d), order="F")
    rc = L.lib().lqg_lambda_pinv(d, m, L.ptr(Lam), L.make_opts(), L.ptr(Lp))
    if rc == 8:
        raise ValueError("Lambda is rank deficient")
    L.check(rc, "lambda_pinv")
    return Lp


def solve_qp_device(D, dvec, A, b, tol=1e-10, max_iter=200, return_iters=False, polish=True):
    """quadprog::solve.QP(D, dvec, t(A), b)$solution (R/lineqGP.R:530, R/lineqAGP.R:431) on the
    GPU: interior-point iteration with every O(n^2 c)/O(n^3) product on the FP64 tensor cores,
    finished by an active-set polish that returns the exact minimiser (polish=False: the
    interior-point iterate)."""
    D = L.f64(D)
    n = D.shape[0]
    dvec = np.ascontiguousarray(np.ravel(dvec), dtype=np.float64)
    A = L.f64(np.asarray(A, float).reshape(-1, n))
    b = np.ascontiguousarray(np.ravel(b), dtype=np.float64)
    x = np.empty(n)
    it = C.c_int32(0)
    rc = L.lib().lqg_solve_qp(n, A.shape[0], L.ptr(D), L.ptr(dvec), L.ptr(A) if A.size else None,
                              L.ptr(b) if b.size else None, float(tol) if polish else -float(tol), int(max_iter), L.make_opts(),
                              L.ptr(x), C.byref(it))
    L.check(rc, "solve_qp")
    return (x, it.value) if return_iters else x


def basis_device(x, ulist):
    """[Phi_1 | ... | Phi_D], Phi_k = basisCompute.lineqGP(x[, k], u_k) (R/lineqGP.R:29-60,
    R/lineqAGP.R:341-348) computed on the GPU; ulist = one knot vector (1-D) or a list of D."""
    if not isinstance(ulist, (list, tuple)):
        ulist = [ulist]
    D = len(ulist)
    x = L.f64(np.asarray(x, float).reshape(-1, D))
    n_t = x.shape[0]
    m_k = np.ascontiguousarray([np.size(u) for u in ulist], dtype=np.int32)
    u = np.ascontiguousarray(np.concatenate([np.ravel(u) for u in ulist]), dtype=np.float64)
    Phi = np.empty((n_t, int(m_k.sum())), order="F")
    rc = L.lib().lqg_basis(n_t, D, L.ptr(m_k), L.ptr(x), n_t, L.ptr(u), L.make_opts(), L.ptr(Phi), n_t)
    if rc == 1 and "knot range" in L.last_error():
        raise ValueError(L.last_error())
    L.check(rc, "basis")
    return Phi
