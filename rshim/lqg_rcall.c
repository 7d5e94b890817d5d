/* .Call shim between R and liblineqgpr_b200.so — SOURCE ONLY in this repository.
 *
 * R (and Rinternals.h) are not installed in the build image, so this file is not compiled
 * or executed here; it shows the reference-side binding a maintainer adds.
 * It replaces the only FFI of the reference,
 *     .Call("rtmg", n+burn.in, seed, initial2, numlin, f2, g2, numquad, quads, PACKAGE="tmg")
 *                                                                       tmg:R/tmg.R:104
 * and adds the box-form / RSM / path entry points used by the patched lineqGPsamplers.R.
 * All matrices arrive column-major, exactly as the C ABI expects; outputs are allocated with
 * Rf_allocMatrix and PROTECTed; status codes become R errors (tmg itself only cat()s and
 * returns NULL, tmg:R/tmg.R:4-57).
 *
 * Build (where R exists):  R CMD SHLIB lqg_rcall.c -I../include -L../lineqgpr_b200 -llineqgpr_b200
 */
#include <R.h>
#include <Rinternals.h>
#include <string.h>

#include "lineqgpr_b200.h"

static void lqg_check(int rc, const char* where) {
  if (rc != LQG_OK) Rf_error("%s: lineqgpr_b200 status %d: %s", where, rc, lqg_last_error());
}

/* drop-in for tmg's native entry: same argument list, same n x d (row = sample) result */
SEXP lqg_rtmg_call(SEXP n_, SEXP seed_, SEXP initial_, SEXP numlin_, SEXP F_, SEXP g_,
                   SEXP numquad_, SEXP quads_) {
  const int n = Rf_asInteger(n_), seed = Rf_asInteger(seed_);
  const int numlin = Rf_asInteger(numlin_), numquad = Rf_asInteger(numquad_);
  const int d = Rf_length(initial_);
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, n, d));
  int rc = lqg_rtmg(n, seed, REAL(initial_), d, numlin, numlin > 0 ? REAL(F_) : NULL,
                    numlin > 0 ? REAL(g_) : NULL, numquad, numquad > 0 ? REAL(quads_) : NULL,
                    REAL(out));
  UNPROTECT(1);
  lqg_check(rc, "lqg_rtmg");
  return out;
}

/* tmvrnorm.HMC body: returns the d x nsim matrix directly (R/lineqGPsamplers.R:142) */
SEXP lqg_hmc_box_call(SEXP mu_, SEXP U_, SEXP Sigma_, SEXP lb_, SEXP ub_, SEXP init_,
                      SEXP n_per_chain_, SEXP burn_in_, SEXP n_chains_, SEXP seed_, SEXP fork_) {
  const int d = Rf_length(mu_);
  const int n_per = Rf_asInteger(n_per_chain_), burn = Rf_asInteger(burn_in_);
  lqg_opts o;
  memset(&o, 0, sizeof(o));
  o.mem = LQG_MEM_HOST;
  o.n_chains = Rf_asInteger(n_chains_);
  o.seed = (uint64_t)Rf_asInteger(seed_);
  o.prepass = -1;
  if (Rf_length(fork_) >= 2) { o.fork = INTEGER(fork_)[0]; o.fork_burn_in = INTEGER(fork_)[1]; }   /* c(fork, fork.burn.in) */
  lqg_hmc_stats st;
  /* Rf_allocMatrix takes int dimensions: n.chains * n.per is formed in 64 bits and refused when it
   * does not fit (the element count d * ncol itself may exceed 2^31: R long vectors hold it) */
  const long long ncol = (long long)o.n_chains * (long long)n_per;
  if (o.n_chains <= 0 || n_per <= 0 || ncol > 2147483647LL)
    Rf_error("lqg_hmc_box: n.chains * n.per = %lld columns do not fit an R matrix dimension", ncol);
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, d, (int)ncol));
  int rc = lqg_hmc_box(d, REAL(mu_), Rf_isNull(U_) ? NULL : REAL(U_), 1, REAL(Sigma_), REAL(lb_), REAL(ub_), REAL(init_), 0,
                       n_per, burn, &o, REAL(out), NULL, &st);
  UNPROTECT(1);
  lqg_check(rc, "lqg_hmc_box");
  if (st.violations != 0) Rf_error("lqg_hmc_box: %lld emitted values outside the bounds", (long long)st.violations);
  return out;
}

/* chains lqg_hmc_box keeps resident for dimension d: the default of control$n.chains */
SEXP lqg_hmc_box_slots_call(SEXP d_) {
  return Rf_ScalarInteger(lqg_hmc_box_slots(Rf_asInteger(d_)));
}

/* The tail of simulate.lineqGP / simulate.lineqAGP in ONE call (R/lineqGP.R:611-640, R/lineqAGP.R:558-575):
 * tmvrnorm.HMC -> xi.sim = qr.solve(Lambda, eta) -> ysim = Phi.test %*% xi.sim -> bands, with eta, xi and
 * ysim resident on the GPU.  Phi_ may be NULL (lineqAGP returns xiAll.sim only); keep_ysim = FALSE never
 * materialises ysim.  Returns list(xi.sim, ysim | NULL, mean | NULL, qtls | NULL). */
SEXP lqg_simulate_call(SEXP mu_, SEXP Sigma_, SEXP lb_, SEXP ub_, SEXP init_, SEXP Lambda_, SEXP Phi_, SEXP probs_,
                       SEXP n_per_chain_, SEXP burn_in_, SEXP n_chains_, SEXP seed_, SEXP keep_ysim_, SEXP fork_) {
  const int d = Rf_length(mu_), m = Rf_ncols(Lambda_);
  const int n_per = Rf_asInteger(n_per_chain_);
  const int have_phi = !Rf_isNull(Phi_), keep = have_phi && Rf_asInteger(keep_ysim_);
  const int n_t = have_phi ? Rf_nrows(Phi_) : 0, np = have_phi ? Rf_length(probs_) : 0;
  if (Rf_nrows(Lambda_) != d) Rf_error("non-conformable arguments");
  if (have_phi && Rf_ncols(Phi_) != m) Rf_error("non-conformable arguments");
  lqg_opts o;
  memset(&o, 0, sizeof(o));
  o.mem = LQG_MEM_HOST;
  o.n_chains = Rf_asInteger(n_chains_);
  o.seed = (uint64_t)Rf_asInteger(seed_);
  o.prepass = -1;
  if (Rf_length(fork_) >= 2) { o.fork = INTEGER(fork_)[0]; o.fork_burn_in = INTEGER(fork_)[1]; }   /* c(fork, fork.burn.in) */
  const long long ncol = (long long)o.n_chains * (long long)n_per;
  if (o.n_chains <= 0 || n_per <= 0 || ncol > 2147483647LL)
    Rf_error("lqg_simulate: n.chains * n.per = %lld columns do not fit an R matrix dimension", ncol);
  SEXP xi = PROTECT(Rf_allocMatrix(REALSXP, m, (int)ncol));
  SEXP ysim = PROTECT(keep ? Rf_allocMatrix(REALSXP, n_t, (int)ncol) : R_NilValue);
  SEXP mean = PROTECT(have_phi ? Rf_allocVector(REALSXP, n_t) : R_NilValue);
  SEXP q = PROTECT(have_phi ? Rf_allocMatrix(REALSXP, np, n_t) : R_NilValue);
  lqg_simulate_args a;
  memset(&a, 0, sizeof(a));
  a.d = d; a.m = m;
  a.mu = REAL(mu_); a.Sigma = REAL(Sigma_); a.lb = REAL(lb_); a.ub = REAL(ub_); a.init = REAL(init_);
  a.Lambda = REAL(Lambda_);
  a.n_t = n_t;
  if (have_phi) { a.Phi = REAL(Phi_); a.ldphi = n_t; a.probs = REAL(probs_); a.nprobs = np; a.mean = REAL(mean); a.qtls = REAL(q); }
  if (keep) { a.ysim_out = REAL(ysim); a.ldy = n_t; }
  a.n_per_chain = n_per; a.burn_in = Rf_asInteger(burn_in_);
  a.xi_out = REAL(xi);
  lqg_simulate_stats st;
  int rc = lqg_simulate(&a, &o, NULL, &st);
  SEXP res = PROTECT(Rf_allocVector(VECSXP, 4));
  SET_VECTOR_ELT(res, 0, xi);
  SET_VECTOR_ELT(res, 1, ysim);
  SET_VECTOR_ELT(res, 2, mean);
  SET_VECTOR_ELT(res, 3, q);
  UNPROTECT(5);
  lqg_check(rc, "lqg_simulate");
  if (st.hmc.violations != 0) Rf_error("lqg_simulate: %lld emitted values outside the bounds", (long long)st.hmc.violations);
  return res;
}

/* tmvrnorm.RSM body */
SEXP lqg_rsm_call(SEXP map_, SEXP factor_, SEXP w_, SEXP lb_, SEXP ub_, SEXP nsim_, SEXP seed_) {
  const int d = Rf_length(map_);
  const int nsim = Rf_asInteger(nsim_);
  lqg_opts o;
  memset(&o, 0, sizeof(o));
  o.mem = LQG_MEM_HOST;
  o.seed = (uint64_t)Rf_asInteger(seed_);
  lqg_rsm_stats st;
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, d, nsim));
  int rc = lqg_rsm(d, REAL(map_), REAL(factor_), REAL(w_), REAL(lb_), REAL(ub_), nsim, 0, &o, REAL(out), &st);
  UNPROTECT(1);
  lqg_check(rc, "lqg_rsm");
  return out;
}

/* accept-reject loop of TruncatedNormal::mvrandn: returns the d x nsim accepted proposals Z */
SEXP lqg_expt_call(SEXP L0_, SEXP l_, SEXP u_, SEXP mu_, SEXP psistar_, SEXP nsim_, SEXP seed_) {
  const int d = Rf_length(l_);
  const int nsim = Rf_asInteger(nsim_);
  lqg_opts o;
  memset(&o, 0, sizeof(o));
  o.mem = LQG_MEM_HOST;
  o.seed = (uint64_t)Rf_asInteger(seed_);
  lqg_expt_stats st;
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, d, nsim));
  int rc = lqg_expt(d, REAL(L0_), REAL(l_), REAL(u_), REAL(mu_), Rf_asReal(psistar_), nsim, 0, &o, REAL(out), &st);
  UNPROTECT(1);
  lqg_check(rc, "lqg_expt");
  return out;
}

/* set-up of TruncatedNormal::mvrandn on the device (cholperm, nleq, psy):
 * list(L0, l, u, mu, psistar, perm (1-based), Lfull) */
SEXP lqg_expt_setup_call(SEXP Sigma_, SEXP l_, SEXP u_) {
  const int d = Rf_length(l_);
  if (Rf_nrows(Sigma_) != d || Rf_ncols(Sigma_) != d || Rf_length(u_) != d) Rf_error("l, u, and Sig have to match in dimension with u>l");
  lqg_opts o;
  memset(&o, 0, sizeof(o));
  o.mem = LQG_MEM_HOST;
  SEXP L0 = PROTECT(Rf_allocMatrix(REALSXP, d, d));
  SEXP lo = PROTECT(Rf_allocVector(REALSXP, d));
  SEXP uo = PROTECT(Rf_allocVector(REALSXP, d));
  SEXP mu = PROTECT(Rf_allocVector(REALSXP, d));
  SEXP ps = PROTECT(Rf_allocVector(REALSXP, 1));
  SEXP perm = PROTECT(Rf_allocVector(INTSXP, d));
  SEXP Lf = PROTECT(Rf_allocMatrix(REALSXP, d, d));
  int32_t iters = 0;
  int rc = lqg_expt_setup(d, REAL(Sigma_), REAL(l_), REAL(u_), &o, REAL(L0), REAL(lo), REAL(uo), REAL(mu), NULL, REAL(ps),
                          INTEGER(perm), REAL(Lf), &iters);
  for (int i = 0; i < d; ++i) INTEGER(perm)[i] += 1;           /* R indices */
  SEXP res = PROTECT(Rf_allocVector(VECSXP, 7));
  SET_VECTOR_ELT(res, 0, L0); SET_VECTOR_ELT(res, 1, lo); SET_VECTOR_ELT(res, 2, uo); SET_VECTOR_ELT(res, 3, mu);
  SET_VECTOR_ELT(res, 4, ps); SET_VECTOR_ELT(res, 5, perm); SET_VECTOR_ELT(res, 6, Lf);
  UNPROTECT(8);
  lqg_check(rc, "lqg_expt_setup");
  return res;
}

/* A %*% B on the FP64 tensor cores: xi.sim = Lambda^+ eta, ysim = Phi.test %*% xi.sim */
SEXP lqg_dgemm_call(SEXP A_, SEXP B_) {
  const int m = Rf_nrows(A_), k = Rf_ncols(A_), n = Rf_ncols(B_);
  if (Rf_nrows(B_) != k) Rf_error("non-conformable arguments");
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, m, n));
  int rc = lqg_dgemm(m, n, k, REAL(A_), m, REAL(B_), k, REAL(out), m, NULL, NULL, NULL);
  UNPROTECT(1);
  lqg_check(rc, "lqg_dgemm");
  return out;
}

/* list(mean = apply(ysim,1,mean), qtls = apply(ysim,1,quantile,probs)) */
SEXP lqg_bands_call(SEXP ysim_, SEXP probs_) {
  const int n_t = Rf_nrows(ysim_), np = Rf_length(probs_);
  const long long N = Rf_ncols(ysim_);
  SEXP mean = PROTECT(Rf_allocVector(REALSXP, n_t));
  SEXP q = PROTECT(Rf_allocMatrix(REALSXP, np, n_t));
  int rc = lqg_bands(n_t, N, REAL(ysim_), n_t, REAL(probs_), np, REAL(mean), REAL(q), NULL, NULL);
  SEXP res = PROTECT(Rf_allocVector(VECSXP, 2));
  SET_VECTOR_ELT(res, 0, mean);
  SET_VECTOR_ELT(res, 1, q);
  UNPROTECT(3);
  lqg_check(rc, "lqg_bands");
  return res;
}

/* ---- set-up stage ------------------------------------- */

/* Ri of tmg:R/tmg.R:27 for M = solve(Sigma): upper U with U %*% t(U) == Sigma */
SEXP lqg_whiten_box_call(SEXP Sigma_) {
  const int d = Rf_nrows(Sigma_);
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, d, d));
  int rc = lqg_whiten_box(d, REAL(Sigma_), NULL, REAL(out), NULL);
  UNPROTECT(1);
  lqg_check(rc, "lqg_whiten_box");                  /* LQG_ERR_NOT_PD: "Error: M must be positive definite." */
  return out;
}

/* min(eigen(S)$values) > 0   (tmg:R/tmg.R:17, R/lineqGP.R:527) */
SEXP lqg_is_pd_call(SEXP S_) {
  int bad = 0;
  int rc = lqg_chol(Rf_nrows(S_), REAL(S_), NULL, NULL, &bad);
  if (rc != LQG_OK && rc != LQG_ERR_NOT_PD) lqg_check(rc, "lqg_chol");
  return Rf_ScalarLogical(rc == LQG_OK);
}

/* solve(A, B), A symmetric positive definite  (chol2inv(chol(A)) %*% B) */
SEXP lqg_solve_spd_call(SEXP A_, SEXP B_) {
  const int d = Rf_nrows(A_), nrhs = Rf_isMatrix(B_) ? Rf_ncols(B_) : 1;
  SEXP out = PROTECT(Rf_isMatrix(B_) ? Rf_allocMatrix(REALSXP, d, nrhs) : Rf_allocVector(REALSXP, d));
  int rc = lqg_solve_spd(d, nrhs, REAL(A_), REAL(B_), NULL, REAL(out));
  UNPROTECT(1);
  lqg_check(rc, "lqg_solve_spd");
  return out;
}

/* list(eta.mu, eta.map, Sigma.eta) of R/lineqGP.R:611-614 */
SEXP lqg_eta_params_call(SEXP Lambda_, SEXP mu_, SEXP map_, SEXP Sigma_, SEXP nugget_) {
  const int d = Rf_nrows(Lambda_), m = Rf_ncols(Lambda_);
  SEXP emu = PROTECT(Rf_allocVector(REALSXP, d));
  SEXP emap = PROTECT(Rf_allocVector(REALSXP, d));
  SEXP Se = PROTECT(Rf_allocMatrix(REALSXP, d, d));
  int rc = lqg_eta_params(d, m, REAL(Lambda_), REAL(mu_), REAL(map_), REAL(Sigma_), Rf_asReal(nugget_), NULL,
                          REAL(emu), REAL(emap), REAL(Se));
  SEXP res = PROTECT(Rf_allocVector(VECSXP, 3));
  SET_VECTOR_ELT(res, 0, emu);
  SET_VECTOR_ELT(res, 1, emap);
  SET_VECTOR_ELT(res, 2, Se);
  UNPROTECT(4);
  lqg_check(rc, "lqg_eta_params");
  return res;
}

/* the m x d operator of qr.solve(Lambda, .)  (R/lineqGP.R:639) */
SEXP lqg_lambda_pinv_call(SEXP Lambda_) {
  const int d = Rf_nrows(Lambda_), m = Rf_ncols(Lambda_);
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, m, d));
  int rc = lqg_lambda_pinv(d, m, REAL(Lambda_), NULL, REAL(out));
  UNPROTECT(1);
  lqg_check(rc, "lqg_lambda_pinv");
  return out;
}

/* solve.QP(Dmat, dvec, Amat = t(A), bvec)$solution with A passed untransposed (rows = constraints) */
SEXP lqg_solve_qp_call(SEXP D_, SEXP dvec_, SEXP A_, SEXP b_) {
  const int n = Rf_nrows(D_), mc = Rf_length(b_);
  SEXP out = PROTECT(Rf_allocVector(REALSXP, n));
  int rc = lqg_solve_qp(n, mc, REAL(D_), REAL(dvec_), mc ? REAL(A_) : NULL, mc ? REAL(b_) : NULL, 0.0, 0, NULL,
                        REAL(out), NULL);
  UNPROTECT(1);
  lqg_check(rc, "lqg_solve_qp");
  return out;
}

/* basisCompute.lineqGP(x, u) for a 1-D knot vector u (R/lineqGP.R:29-60) */
SEXP lqg_basis_call(SEXP x_, SEXP u_) {
  const int n_t = Rf_length(x_);
  const int32_t m = Rf_length(u_);
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, n_t, m));
  int rc = lqg_basis(n_t, 1, &m, REAL(x_), n_t, REAL(u_), NULL, REAL(out), n_t);
  UNPROTECT(1);
  lqg_check(rc, "lqg_basis");
  return out;
}

/* simulate tail + bands in one call: list(ysim (or NULL), mean, qtls) with ysim = Phi.test %*% xi.sim
 * (R/lineqGP.R:640), mean = apply(ysim,1,mean), qtls = apply(ysim,1,quantile,probs) (R/lineqGPutils.R:422-423);
 * keep_ysim = FALSE never materialises the n_t x N matrix */
SEXP lqg_paths_call(SEXP Phi_, SEXP xi_, SEXP probs_, SEXP keep_ysim_) {
  const int n_t = Rf_nrows(Phi_), m = Rf_ncols(Phi_), np = Rf_length(probs_);
  const long long N = Rf_ncols(xi_);
  if (Rf_nrows(xi_) != m) Rf_error("non-conformable arguments");
  const int keep = Rf_asInteger(keep_ysim_);
  SEXP ysim = PROTECT(keep ? Rf_allocMatrix(REALSXP, n_t, (int)N) : Rf_allocVector(REALSXP, 0));
  SEXP mean = PROTECT(Rf_allocVector(REALSXP, n_t));
  SEXP q = PROTECT(Rf_allocMatrix(REALSXP, np, n_t));
  int rc = lqg_paths(n_t, m, N, REAL(Phi_), n_t, REAL(xi_), m, REAL(probs_), np, keep ? REAL(ysim) : NULL, n_t,
                     REAL(mean), REAL(q), 0, NULL, NULL, NULL);
  SEXP res = PROTECT(Rf_allocVector(VECSXP, 3));
  SET_VECTOR_ELT(res, 0, ysim);
  SET_VECTOR_ELT(res, 1, mean);
  SET_VECTOR_ELT(res, 2, q);
  UNPROTECT(4);
  lqg_check(rc, "lqg_paths");
  return res;
}
