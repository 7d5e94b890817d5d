/* lineqgpr_b200 — C ABI of the B200-native constrained-posterior simulation path.
 *
 * This header is the drop-in boundary.  Plain C: pointers, sizes and
 * status codes only; no torch / C++ types.  All matrices are FP64, COLUMN-major, exactly
 * as R lays them out.  Reference citations:
 *   R/...   = /root/reference/dev/lineqGPR/R/...
 *   tmg:... = file inside /root/reference/dependecies/tmg_0.3.tar.gz
 *
 * Every entry point returns an lqg status code and never throws / exits across the ABI
 * (tmg's C++ has no error path at all and can loop forever, tmg:src/HmcSampler.cpp:51).
 * There is NO CPU fallback: without a CUDA device every compute entry point returns
 * LQG_ERR_CUDA.
 */
#ifndef LINEQGPR_B200_H
#define LINEQGPR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ------------------------------------------------------------------- */
#define LQG_OK                   0
#define LQG_ERR_INVALID          1   /* bad argument (sizes, NULL, l >= u ...)                 */
#define LQG_ERR_UNSUPPORTED      2   /* quadratic constraints (numquad > 0): dead path in
                                        lineqGPR (R/lineqGPsamplers.R:140-141 passes q=NULL)  */
#define LQG_ERR_INIT_INFEASIBLE  3   /* tmg:R/tmg.R:44-47 "initial point violates ..."        */
#define LQG_ERR_CUDA             4   /* no device / CUDA runtime error (see lqg_last_error)   */
#define LQG_ERR_RESTART_CAP      5   /* a chain exceeded max_restarts in one transition        */
#define LQG_ERR_DRAWS_EXHAUSTED  6   /* injected Gaussian stream too short                     */
#define LQG_ERR_INTERNAL         9   /* host allocation failure or another internal exception  */
#define LQG_ERR_NOT_PD           8   /* tmg:R/tmg.R:17-20 "M must be positive definite"        */
#define LQG_ERR_BUDGET           7   /* RSM: proposal budget exhausted before nsim accepts     */

/* where the caller's pointers live */
#define LQG_MEM_HOST    0
#define LQG_MEM_DEVICE  1

/* ---- options shared by the samplers ------------------------------------------------- */
typedef struct lqg_opts {
  int32_t  mem;            /* LQG_MEM_HOST: all pointers are host memory, the call copies in/out
                              (this is what the .Call shim uses); LQG_MEM_DEVICE: all array
                              pointers are device memory on the current device                */
  int32_t  hmc_search;     /* lqg_hmc_box / lqg_simulate hit search: 0 = default (walls ranked by
                              tan(t/2) of their hit time, tmg's own acos/atan2 arithmetic only where a
                              discrete threshold, a near-tie or the end of the trajectory is close),
                              1 = every search with tmg's arithmetic (tmg:src/HmcSampler.cpp:157-181),
                              2 = like 0 but every search looks at the whole remaining time (0 looks at a
                              short time window first and widens it when it holds no hit)             */
  void*    stream;         /* cudaStream_t to launch on (NULL = legacy default stream)         */
  uint64_t seed;           /* Philox key; replaces sample(1:1e7,1) of tmg:R/tmg.R:102          */
  int32_t  n_chains;       /* HMC: independent chains. 1 = the reference's single serial chain */
  int32_t  max_restarts;   /* HMC: cap on velocity/feasibility restarts per transition
                              (0 -> 100000); the reference loops forever                        */
  int64_t  chain_offset;   /* HMC: global id of this call's first chain. RNG streams are keyed
                              by (seed, global chain id), so a job sharded over GPUs gives the
                              same samples as the unsharded job                                 */
  const double* injected;  /* nullable. HMC: [n_chains][n_injected] N(0,1) draws consumed in
                              order, d per (re)start, exactly like nd(eng1) at
                              tmg:src/HmcSampler.cpp:54-55.  RSM: d x n_injected normals,
                              one column per proposal                                           */
  int64_t  n_injected;     /* HMC: draws available per chain; RSM: number of proposal columns  */
  const double* injected_u;/* RSM only, nullable: uniforms, one per IN-BOX proposal
                              (R/lineqGPsamplers.R:55)                                          */
  int64_t  n_injected_u;
  int32_t  prepass;        /* HMC box: 1 = batch the fresh momenta U*a of all transitions as one
                              FP64 tensor-core GEMM before the chain kernel, 0 = per-chain
                              in-kernel product, -1 = auto                                      */
  int32_t  fork;           /* lqg_hmc_box / lqg_simulate: F > 1 = chains share their burn-in in families of F.
                              Chains with global ids [kF, (k+1)F) start from the state that ONE root chain
                              (Philox stream of chain kF) has reached after burn_in transitions from `init`,
                              run fork_burn_in more transitions each on their own streams, then emit: the
                              burn-in work is burn_in * n_chains / F + fork_burn_in * n_chains instead of
                              burn_in * n_chains, every emitted column still comes from a chain that is
                              burn_in + fork_burn_in transitions away from the start.  Samples depend on
                              (seed, global id, F), not on the sharding.  0 / 1 = independent chains       */
  int32_t  fork_burn_in;   /* transitions a forked chain runs before it emits (see fork)                  */
  int32_t  canon_mode;     /* lqg_hmc_canonical: 0 = auto, 1 = one persistent CTA per chain (F in shared
                              memory when it fits), 2 = lock-step rounds, the dot products of all
                              chains as one FP64 tensor-core GEMM per round (large dense F)        */
} lqg_opts;

typedef struct lqg_hmc_stats {
  int64_t transitions;     /* completed transitions, burn-in included                          */
  int64_t bounces;         /* wall reflections (tmg:src/HmcSampler.cpp:90-112 executions)      */
  int64_t hit_searches;    /* _getNextLinearHitTime calls (tmg:src/HmcSampler.cpp:152)         */
  int64_t vel_restarts;    /* velsign < 0 restarts (tmg:src/HmcSampler.cpp:112,117)            */
  int64_t chk_restarts;    /* end-point infeasible restarts (tmg:src/HmcSampler.cpp:121-130)   */
  int64_t exact_evals;     /* constraints that survived the filter and had their hit time evaluated */
  int64_t violations;      /* emitted samples outside the constraints: always 0                */
  int64_t failed_chains;   /* chains stopped by max_restarts / exhausted injected draws        */
  double  chain_ms;        /* device time of the chain kernel(s)                               */
  double  prepass_ms;      /* device time of the momentum GEMM pre-pass (0 when off)           */
  int64_t slow_searches;   /* box kernel: hit searches that were re-run with tmg's exact arithmetic */
  int64_t window_retries;  /* box kernel: time windows without a hit, searched again 4 x wider          */
  /* forked families (lqg_opts.fork > 1): the share of the totals above spent in the root phase (one chain per
   * family burning in; at most one chain per SM runs then, so it is latency- rather than throughput-bound)   */
  int64_t root_transitions, root_bounces, root_hit_searches, root_restarts;
  double  root_chain_ms, root_prepass_ms;
} lqg_hmc_stats;

typedef struct lqg_rsm_stats {
  int64_t proposals;       /* mvrnorm draws consumed (R/lineqGPsamplers.R:50,52)               */
  int64_t in_box;          /* proposals inside [lb,ub] (R/lineqGPsamplers.R:51)                */
  int64_t accepted;        /* u <= t accepts (R/lineqGPsamplers.R:56)                          */
  int64_t rounds;          /* kernel rounds                                                    */
  double  kernel_ms;
} lqg_rsm_stats;

/* ---- HMC ---------------------------------------------------------------------------- */

/* Replaces the native entry of tmg:
 *   RcppExport SEXP rtmg(SEXP n_, SEXP seed_, SEXP initial_, SEXP numlin_, SEXP F_, SEXP g_,
 *                        SEXP numquad_, SEXP quadratics_)      tmg:src/rtmg.h:6, rtmg.cpp:14-66
 * called as .Call("rtmg", n+burn.in, seed, initial2, numlin, f2, g2, numquad, quads)
 *                                                              tmg:R/tmg.R:104
 * Same contract: canonical (whitened) frame, one chain, HOST pointers, F is numlin x d
 * column-major, out is n x d column-major with ROW = sample (all n rows returned, the
 * caller drops burn-in, tmg:R/tmg.R:105).  numquad must be 0.
 * Momentum draws are Philox-4x32-10 keyed by `seed` (not libstdc++'s knuth_b): the chain is
 * statistically, not bitwise, the reference's chain. */
int lqg_rtmg(int32_t n, int32_t seed, const double* initial, int32_t d, int32_t numlin,
             const double* F, const double* g, int32_t numquad, const double* quadratics,
             double* out);

/* General dense constraints F z + g >= 0 in the canonical frame, many chains.
 * Transition = tmg:src/HmcSampler.cpp:43-133 (sampleNext), hit search :152-189,
 * feasibility :244-265.
 *   F        c x d column-major; g[c]
 *   initial  d x n_chains column-major when initial_ld >= d, one shared d-vector when
 *            initial_ld == 0 (tmg:src/rtmg.cpp:55 setInitialValue)
 *   out      d x (n_chains * n_per_chain) column-major, COLUMN = sample, chain-major:
 *            column chain*n_per_chain + s is the (burn_in + s)-th transition of `chain`
 *   state_out nullable, d x n_chains: last state of every chain (to continue a run)
 */
int lqg_hmc_canonical(int32_t d, int32_t c, const double* F, const double* g,
                      const double* initial, int64_t initial_ld,
                      int32_t n_per_chain, int32_t burn_in, const lqg_opts* opts,
                      double* out, double* state_out, lqg_hmc_stats* stats);

/* Parity tooling: lqg_hmc_canonical for n_per_chain transitions (burn_in = 0) that also records,
 * for every chain, the first trace_cap wall hits as (hit time t, constraint row cn) — the pair
 * _getNextLinearHitTime returns (tmg:src/HmcSampler.cpp:152-189) — in trace_t / trace_cn
 * [n_chains x trace_cap, chain-major]; trace_n[n_chains] = number of hits seen.  Host pointers
 * only.  Used to teacher-force single hit searches against the CPU oracle's trace. */
int lqg_hmc_canonical_trace(int32_t d, int32_t c, const double* F, const double* g,
                            const double* initial, int64_t initial_ld, int32_t n_per_chain,
                            const lqg_opts* opts, double* out, int32_t trace_cap,
                            double* trace_t, int32_t* trace_cn, int32_t* trace_n);

/* Parity tooling: the two device-side hit-time functions over n independent (f.a, f.b, g) triples
 * (host pointers).  t_exact[n] (nullable) = tmg's hit time of one constraint,
 * tmg:src/HmcSampler.cpp:157-181, 0 = none — compared with the oracle's scalar restatement on
 * randomised inputs.  fast_class[n] (nullable) = verdict of the trig-free ranking the box kernel
 * uses (0 = cannot be hit within tan(t/2) <= klim, 1 = hit: fast_key = tan(t/2), fast_sin / fast_cos
 * of that t, 2 = too close to one of tmg's thresholds: decided by the exact function).            */
int lqg_hit_time_batch(int64_t n, const double* fa, const double* fb, const double* g, double klim,
                       double* t_exact, int32_t* fast_class, double* fast_key, double* fast_sin, double* fast_cos);

/* Box-constrained form: what tmvrnorm.HMC always builds (R/lineqGPsamplers.R:128-141:
 * f = [I; -I] rows with finite bounds, g = c(-lb, ub)), sampled directly in the eta frame.
 * With U = Ri of tmg:R/tmg.R:27 (so U U' = Sigma) the chain is algebraically the reference's
 * whitened chain: x_a = U a, a bounce on coordinate j reflects with
 * column j of Sigma, the emitted sample is mu + x_b and needs no un-whitening
 * (tmg:R/tmg.R:106).
 *   mu[d], lb[d], ub[d] (+-Inf allowed), init[d] or d x n_chains (init_ld as above): the
 *   clamped start of R/lineqGPsamplers.R:133-135 in the ORIGINAL eta frame
 *   U d x d column-major; u_upper != 0 promises U is upper triangular (tmg's Ri is)
 *   Sigma d x d column-major
 *   out d x (n_chains*n_per_chain), COLUMN = sample (the d x nsim layout tmvrnorm returns,
 *   R/lineqGPsamplers.R:142)
 * Returns LQG_ERR_INIT_INFEASIBLE unless lb < init < ub strictly (tmg:R/tmg.R:44-47).
 */
int lqg_hmc_box(int32_t d, const double* mu, const double* U, int32_t u_upper,
                const double* Sigma, const double* lb, const double* ub,
                const double* init, int64_t init_ld,
                int32_t n_per_chain, int32_t burn_in, const lqg_opts* opts,
                double* out, double* state_out, lqg_hmc_stats* stats);

/* Number of chains lqg_hmc_box keeps resident at once for dimension d on the current device
 * (SMs x CTAs per SM of the kernel variant it would launch).  n_chains equal to a multiple of
 * this fills every wave; the host wrappers use it as the default chain count.  <= 0 on error. */
int lqg_hmc_box_slots(int32_t d);

/* ---- RSM ---------------------------------------------------------------------------- */

/* Rejection sampling from the mode, R/lineqGPsamplers.R:36-61.
 *   map[d]; factor d x d column-major with x = map + factor z (MASS::mvrnorm's eigen factor,
 *   R/lineqGPsamplers.R:50); w[d] = Sigma^-1 map (R/lineqGPsamplers.R:42-43); lb, ub [d]
 *   out d x nsim, the first nsim accepted proposals in proposal order
 *   max_proposals: budget (0 -> 2^40)
 */
int lqg_rsm(int32_t d, const double* map, const double* factor, const double* w,
            const double* lb, const double* ub, int64_t nsim, int64_t max_proposals,
            const lqg_opts* opts, double* out, lqg_rsm_stats* stats);

/* ---- ExpT --------------------------------------------------------------------------- */

typedef struct lqg_expt_stats {
  int64_t proposals;       /* tilted proposals consumed up to the nsim-th accept                */
  int64_t accepted;
  int64_t rounds;
  double  kernel_ms;
} lqg_expt_stats;

/* The accept-reject loop of TruncatedNormal::mvrandn(l, u, Sig, n), i.e. the per-sample part of
 * tmvrnorm.ExpT (R/lineqGPsamplers.R:179-185); TruncatedNormal is not vendored in the reference,
 * this follows Botev (2017), JRSS-B 79(1):125-148.  The host computes (as the R package does)
 * the permuted Cholesky factor, its scaling and the tilt:
 *   L0    d x d column-major: row-scaled factor with ZERO diagonal
 *   l, u  scaled and permuted bounds; mu tilt parameters with mu[d-1] = 0; psistar = psi(x*, mu*)
 * Z_out (d x nsim): the first nsim ACCEPTED proposals Z in proposal order; the caller finishes
 * with  x = Lfull[order,] %*% Z  (lqg_dgemm).  max_proposals: budget (0 -> 2^40). */
int lqg_expt(int32_t d, const double* L0, const double* l, const double* u, const double* mu,
             double psistar, int64_t nsim, int64_t max_proposals, const lqg_opts* opts,
             double* Z_out, lqg_expt_stats* stats);

/* Set-up of tmvrnorm.ExpT on the device: what TruncatedNormal::mvrandn (reached from
 * R/lineqGPsamplers.R:179-185) does before its accept-reject loop — the variable-reordered Cholesky
 * factor of Sigma (Gibson, Glasbey & Elston) and the Newton solve for the minimax tilt (Botev 2017,
 * eq. 11).  l, u = lb - mu, ub - mu [d].  Outputs (each [d] unless noted): L0 d x d (row-scaled factor,
 * zero diagonal), l_out / u_out (permuted, scaled), mu_out (tilt, last entry 0), x_out (last entry 0),
 * *psistar, perm (int32: variable perm[j] comes j-th), Lfull d x d (unscaled factor) — the arguments of
 * lqg_expt and of the back-transformation  x = Lfull[order, ] %*% Z.  *iters = Newton iterations.
 * LQG_ERR_NOT_PD: Sigma is not positive semi-definite; LQG_ERR_INVALID: the Newton iteration failed
 * ("Covariance matrix is ill-conditioned and method failed").  d <= 2048.                          */
int lqg_expt_setup(int32_t d, const double* Sigma, const double* l, const double* u, const lqg_opts* opts,
                   double* L0, double* l_out, double* u_out, double* mu_out, double* x_out, double* psistar,
                   int32_t* perm, double* Lfull, int32_t* iters);

/* ---- sample paths and bands --------------------------------------------------------- */

/* C (m x n) = A (m x k) * B (k x n), all column-major FP64, on the FP64 tensor cores.
 * This is xi.sim = Lambda^+ eta and ysim = Phi.test %*% xi.sim of R/lineqGP.R:639-640 and the
 * user-side PhiAll.test %*% xiAll.sim of R/lineqAGP.R:509-512.
 * row_sum nullable [m]: when given, sum_j C[i,j] is accumulated (band mean numerator);
 * C nullable when only row sums are wanted (config 5: ysim does not fit anywhere). */
int lqg_dgemm(int32_t m, int32_t n, int32_t k, const double* A, int64_t lda,
              const double* B, int64_t ldb, double* C, int64_t ldc, double* row_sum,
              const lqg_opts* opts, double* kernel_ms);

/* Row-wise mean and R type-7 quantiles of ysim (n_t x N, column-major):
 * apply(ysim, 1, mean); apply(ysim, 1, quantile, probs)   R/lineqGPutils.R:422-423, :294-296
 *   mean[n_t]; qtls nprobs x n_t column-major (the layout apply() returns) */
int lqg_bands(int32_t n_t, int64_t N, const double* ysim, int64_t ld,
              const double* probs, int32_t nprobs, double* mean, double* qtls,
              const lqg_opts* opts, double* kernel_ms);

/* Sample paths and their bands in one call:
 *   ysim = Phi.test %*% xi.sim  (R/lineqGP.R:640, R/lineqAGP.R:509-512), then
 *   apply(ysim, 1, mean) and apply(ysim, 1, quantile, probs)  (R/lineqGPutils.R:422-423).
 * Phi n_t x m (ldphi >= n_t), xi m x N (ldxi >= m), column-major.  ysim is nullable: the n_t x N
 * matrix is produced slab by slab on the device (row slabs of Phi of at most slab_rows rows,
 * 0 = as many as fit ~4 GB), summarised exactly and dropped, so it never has to exist anywhere
 * (config 5: 10^5 x 10^6 doubles); when ysim (ldy >= n_t) is given the slabs are also copied out.
 * Without ysim, a dense slab with N >= 8192 is not even stored on the device: the product's epilogue
 * counts, collects and sums for the bands (same exact order statistics; ldphi and m even, else the
 * slab is stored and read again).
 * mean[n_t]; qtls nprobs x n_t (nullable when nprobs = 0).  gemm_ms / bands_ms nullable.          */
int lqg_paths(int32_t n_t, int32_t m, int64_t N, const double* Phi, int64_t ldphi,
              const double* xi, int64_t ldxi, const double* probs, int32_t nprobs,
              double* ysim, int64_t ldy, double* mean, double* qtls, int32_t slab_rows,
              const lqg_opts* opts, double* gemm_ms, double* bands_ms);

/* ---- set-up stage on the device --------------------------- */
/* All matrices column-major FP64; pointers host or device according to opts->mem; every
 * product runs on the FP64 tensor cores (blocked algorithms, lqg_linalg.cu).                 */

/* U (d x d, upper triangular, positive diagonal) with U U' = (Sigma + Sigma')/2: the matrix
 * tmg::rtmg whitens with when tmvrnorm.HMC calls it with M = solve(Sigma)
 * (R/lineqGPsamplers.R:139-141; tmg:R/tmg.R:15 symmetrise, :17-20 PD test, :21 R = chol(M),
 * :27 Ri = solve(R)): Ri is upper triangular with Ri Ri' = M^-1 = Sigma, and that factor is
 * unique, so it is computed directly as a reversed Cholesky factorisation — no inverse of the
 * ill-conditioned Sigma.  LQG_ERR_NOT_PD (tmg's "M must be positive definite") otherwise.
 * lqg_hmc_box calls this internally when its U argument is NULL.                              */
int lqg_whiten_box(int32_t d, const double* Sigma, const lqg_opts* opts, double* U_out, double* kernel_ms);

/* chol(): lower factor L with L L' = (A + A')/2.  L_out nullable (test only).  *bad_pivot
 * (nullable) = 0 when A is positive definite, else the 1-based index of the first non-positive
 * pivot and the call returns LQG_ERR_NOT_PD.  Replaces the eigen()-based tests of
 * tmg:R/tmg.R:17 and R/lineqGP.R:527-528 (min eigenvalue <= 0  <=>  the factorisation fails). */
int lqg_chol(int32_t d, const double* A, const lqg_opts* opts, double* L_out, int32_t* bad_pivot);

/* X = solve(A, B) for a symmetric positive definite A (d x d) and B (d x nrhs): the
 * chol2inv/solve products of the posterior algebra (R/lineqGP.R:514-523).                     */
int lqg_solve_spd(int32_t d, int32_t nrhs, const double* A, const double* B, const lqg_opts* opts,
                  double* X_out);

/* eta-space parameters of simulate.lineqGP (R/lineqGP.R:611-614; R/lineqAGP.R:558-566):
 *   eta_mu = Lambda mu, eta_map = Lambda xi_map, Sigma_eta = Lambda Sigma Lambda' + nugget I
 * Lambda d x m, mu[m], xi_map[m], Sigma m x m; outputs eta_mu[d], eta_map[d], Sigma_eta d x d
 * (mu / xi_map / eta_mu / eta_map nullable).                                                  */
int lqg_eta_params(int32_t d, int32_t m, const double* Lambda, const double* mu, const double* xi_map,
                   const double* Sigma, double nugget, const lqg_opts* opts,
                   double* eta_mu, double* eta_map, double* Sigma_eta);

/* Hat-basis matrix of an additive model: PhiAll.test = [Phi_1 | ... | Phi_D] with
 * Phi_k = basisCompute.lineqGP(x[, k], u_k)  (R/lineqGP.R:29-60; R/lineqAGP.R:341-348); D = 1 is
 * the 1-D lineqGP basis.  x n_t x D column-major (ldx >= n_t), m_k[D] knots per block (always a
 * HOST array), u the
 * concatenated increasing knot vectors (sum m_k entries); Phi n_t x sum(m_k) column-major
 * (ld >= n_t).  Bit-identical to the R expression.  LQG_ERR_INVALID when a point lies outside
 * its knot range (R fails there on an NA index).                                              */
int lqg_basis(int32_t n_t, int32_t D, const int32_t* m_k, const double* x, int64_t ldx, const double* u,
              const lqg_opts* opts, double* Phi, int64_t ld);

/* Constrained MAP: min 1/2 x'Dx - dvec'x  s.t.  A x >= b, i.e. quadprog::solve.QP(Dmat, dvec,
 * t(A), b)$solution as called at R/lineqGP.R:530 and R/lineqAGP.R:431 (quadprog itself is not
 * vendored in the reference: a Mehrotra predictor-corrector interior-point method replaces its
 * active-set iteration; both return the unique minimiser).  D n x n symmetric positive definite,
 * dvec[n], A mc x n column-major, b[mc] finite; mc = 0 returns solve(D, dvec).
 * tol == 0 -> 1e-10 on the merit max(|rd|/(1e3 scale_d), |rp|/scale_p, mu/scale_p); when the
 * merit stalls at the rounding floor for 5 iterations the best iterate is returned if its merit
 * is below 1e3 tol.  The result is then polished by an active-set solve (guess lam_i > s_i,
 * Schur-complement solve, KKT check, a few adjustments) so that, like quadprog, the exact
 * minimiser is returned; tol < 0 means tolerance |tol| without the polish.
 * max_iter <= 0 -> 200; *iters nullable.
 * LQG_ERR_NOT_PD when D is not positive definite, LQG_ERR_BUDGET when max_iter is exhausted
 * (x_out then holds the last iterate).                                                        */
int lqg_solve_qp(int32_t n, int32_t mc, const double* D, const double* dvec, const double* A,
                 const double* b, double tol, int32_t max_iter, const lqg_opts* opts,
                 double* x_out, int32_t* iters);

/* Lambda^+ (m x d) with Lambda^+ eta = qr.solve(Lambda, eta) (R/lineqGP.R:639,
 * R/lineqAGP.R:575) for a full-column-rank Lambda (d x m, d >= m): normal equations
 * (Lambda'Lambda) X = Lambda' by blocked Cholesky, followed by one step of iterative refinement
 * X += G^-1 Lambda' (I - Lambda X), which restores QR-level accuracy for the constraint
 * matrices lineqGPR builds (cond(Lambda) <~ 1e4).  LQG_ERR_NOT_PD if Lambda is rank deficient.  */
int lqg_lambda_pinv(int32_t d, int32_t m, const double* Lambda, const lqg_opts* opts, double* Lp_out);

/* ---- the whole simulate tail in one call, device resident, optionally over several GPUs ----- */

/* Communicator of a multi-GPU job: one process per GPU, this process's GPU = the current CUDA
 * device.  Chains shard over the ranks with no data-path collective; lqg_simulate then performs
 * the two exchanges of SURVEY.md par.8(e) option A with NCCL all-gathers over NVLink.  NCCL is
 * resolved at run time (libnccl.so.2; LQG_NCCL_LIB overrides), so single-GPU use does not need it.
 *   lqg_comm_unique_id  rank 0 fills 128 bytes that the host program hands to every rank
 *                       (torch.distributed, MPI, a file ...)
 *   lqg_comm_init       collective over the `world` ranks; world == 1 needs no id and no NCCL    */
typedef struct lqg_comm lqg_comm;
int lqg_comm_unique_id(void* id128);
int lqg_comm_init(const void* id128, int32_t rank, int32_t world, lqg_comm** out);
int lqg_comm_rank(const lqg_comm* c);
int lqg_comm_world(const lqg_comm* c);
int lqg_comm_destroy(lqg_comm* c);
/* all-gather of `count` doubles per rank, DEVICE pointers, on `stream` (recv holds world*count
 * doubles and may alias send at recv + rank*count) */
int lqg_comm_allgather(const lqg_comm* c, const double* send, double* recv, int64_t count, void* stream);

/* Arguments of lqg_simulate: the tail of simulate.lineqGP (R/lineqGP.R:611-640) and of
 * simulate.lineqAGP (R/lineqAGP.R:558-575) plus the product the user applies to its result
 * (R/lineqAGP.R:509-512) and the band summaries drawn from it (R/lineqGPutils.R:422-423).
 * Pointers are host or device memory according to opts->mem.  Column-major FP64 throughout.     */
typedef struct lqg_simulate_args {
  int32_t d, m;              /* d = nrow(Lambda) = length(eta), m = ncol(Lambda) = length(xi)        */
  /* tmvPar <- list(mu = eta.mu, map = eta.map, Sigma = Sigma.eta, lb, ub)   R/lineqGP.R:622        */
  const double* mu;          /* [d]  eta.mu (lineqAGP passes eta.map here, R/lineqAGP.R:567)         */
  const double* Sigma;       /* d x d  Lambda Sigma Lambda' + nugget I   R/lineqGP.R:613-614         */
  const double* lb;          /* [d]                                                                  */
  const double* ub;          /* [d]                                                                  */
  const double* init;        /* [d]  eta.map clamped into the box   R/lineqGPsamplers.R:133-135      */
  const double* U;           /* nullable d x d upper factor with U U' = Sigma (tmg's Ri); NULL = the
                                device factors Sigma itself (lqg_whiten_box)                          */
  /* xi.sim <- qr.solve(Lambda, eta)   R/lineqGP.R:639, R/lineqAGP.R:575                             */
  const double* Lambda;      /* d x m; nullable when Lambda_pinv is given, or when d == m and
                                xi = eta is wanted (both NULL)                                        */
  const double* Lambda_pinv; /* nullable m x d: a precomputed lqg_lambda_pinv result                 */
  /* ysim <- Phi.test %*% xi.sim   R/lineqGP.R:640, R/lineqAGP.R:509-512                             */
  int32_t n_t;               /* test points; 0 = stop after xi                                       */
  const double* Phi;         /* n_t x m (ldphi >= n_t), or NULL with the basis description below     */
  int64_t ldphi;
  int32_t basis_D;           /* Phi == NULL: Phi = [Phi_1 | ... | Phi_D], Phi_k = hat basis of        */
  const int32_t* basis_mk;   /*   xtest[, k] on the knots of block k (lqg_basis; R/lineqGP.R:29-60,   */
  const double* xtest;       /*   R/lineqAGP.R:341-348); basis_mk is a HOST array of D counts,        */
  int64_t ldx;               /*   xtest n_t x D (ldx >= n_t), knots the concatenated knot vectors     */
  const double* knots;
  const double* probs;       /* [nprobs] quantile levels (R type 7), nprobs <= 8                      */
  int32_t nprobs;
  int32_t n_per_chain;       /* samples per chain; this rank draws opts->n_chains * n_per_chain       */
  int32_t burn_in;           /* transitions every chain runs before it emits                          */
  int32_t slab_rows;         /* rows of Phi per GEMM + summary slab, 0 = as many as fit ~4 GB         */
  int32_t reserved0;
  /* results (each nullable) */
  double* eta_out;           /* d x N_local, chain-major columns like lqg_hmc_box                     */
  double* xi_out;            /* m x N_local, same column order: this rank's share of xi.sim           */
  double* ysim_out;          /* rows_local x N_total (ldy >= rows_local): this rank's rows of ysim,   */
  int64_t ldy;               /*   columns in job order = [xi_out of rank 0 | xi_out of rank 1 | ...]  */
  double* mean;              /* [n_t] on EVERY rank                                                   */
  double* qtls;              /* nprobs x n_t on every rank                                            */
} lqg_simulate_args;

typedef struct lqg_simulate_stats {
  lqg_hmc_stats hmc;
  double xi_ms;              /* Lambda^+ eta products (device time, side stream)                      */
  double gemm_ms;            /* Phi xi slabs                                                          */
  double bands_ms;           /* mean + quantiles of the slabs                                         */
  double total_ms;           /* whole call on the device (events on the caller's stream)              */
  int64_t n_local, n_total;  /* samples drawn by this rank / by the job                               */
  int32_t row_begin, row_end;/* rows of Phi this rank summarised                                      */
  int32_t rank, world;
  int32_t rounds;            /* units of samples processed = xi products = xi all-gather calls        */
  int32_t reserved0;
} lqg_simulate_stats;

/* sampler -> xi = Lambda^+ eta -> Phi xi -> bands, with eta, xi and ysim never leaving HBM unless an
 * output pointer asks for them.  The box sampler is lqg_hmc_box (same chains, same Philox streams:
 * eta_out equals what lqg_hmc_box returns for the same opts).  While later sampling rounds run, the
 * samples every chain has finished are multiplied by Lambda^+ on a second stream, copied to xi_out
 * and — with a communicator — all-gathered, so that after the last round every rank holds the xi of
 * the whole job (m x N_total; the work proceeds in units of 16 (one GPU) or 32 (several) samples per chain, the same on every
 * rank, and the gathered columns are ordered unit, rank, chain, sample).  Rank r then
 * summarises rows [r ceil(n_t/world), (r+1) ceil(n_t/world)) of Phi over all N_total samples (exact
 * mean and type-7 quantiles) and the band pieces are all-gathered: every rank returns the full
 * mean[n_t], qtls[nprobs x n_t].  Global chain ids are opts->chain_offset + rank * n_chains + local
 * id, so a job gives the same samples however many GPUs it runs on.  comm NULL = one GPU.          */
int lqg_simulate(const lqg_simulate_args* args, const lqg_opts* opts, lqg_comm* comm,
                 lqg_simulate_stats* stats);

/* Page-locked host memory for large result matrices: with it the LQG_MEM_HOST entry points copy
 * samples out at full PCIe speed and concurrently with the sampling rounds; with ordinary
 * (pageable) memory — e.g. an R matrix — the copies are staged and serialised by the driver. */
void* lqg_alloc_host(uint64_t bytes);     /* NULL on failure (see lqg_last_error) */
void  lqg_free_host(void* p);

/* ---- misc --------------------------------------------------------------------------- */
const char* lqg_version(void);
const char* lqg_last_error(void);        /* thread-local text of the last non-OK status       */
int lqg_device_count(void);              /* number of CUDA devices, 0 when none               */
/* measured peaks used as roofline denominators: DFMA and DMMA micro-benchmarks (Tflop/s) */
int lqg_measure_fp64_peaks(double* dfma_tflops, double* dmma_tflops, const lqg_opts* opts);
/* probe: even warps issue DFMA, odd warps DMMA; *tflops = their combined rate, *dfma_share = the
 * DFMA fraction of the flops.  A sum above either peak would mean separate units. */
int lqg_measure_fp64_mixed(double* tflops, double* dfma_share, const lqg_opts* opts);
/* counts kernel launches made by this library since the last reset (bench gpu_launches) */
int64_t lqg_launch_count(int32_t reset);
/* debugging aid: with LQG_GUARD=1 in the environment every internal device buffer carries a
 * canary band on either side, checked when it is released; returns how many were overwritten */
int lqg_guard_violations(void);

#ifdef __cplusplus
}
#endif
#endif /* LINEQGPR_B200_H */
