// Internal launch interfaces between the C-ABI layer (lqg_capi.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lqg {

// ---- box-form HMC (lqg_hmc_box.cu) -----------------------------------------------------
struct BoxParams {
  int d;
  const double* mu;        // [d]
  const double* lb;        // [d]
  const double* ub;        // [d]
  const double* U;         // d x d col-major, U U' = Sigma
  int u_upper;
  const double* Sigma;     // d x d col-major
  const double* glo;       // [d] mu - lb  (kBigG when lb = -Inf)
  const double* ghi;       // [d] ub - mu  (kBigG when ub = +Inf)
  const double* sdiag;     // [d] Sigma_jj
  double abs_eps;          // absolute part of the filter margin
  double* state;           // d x n_chains, centred position x_b, in/out
  int n_chains;
  const int* active;       // nullable: chain ids to run in this launch (NULL = 0..n_active-1)
  int n_active;
  int attempts;            // momenta each active chain may consume in this launch
  const double* pool;      // nullable: pre-computed momenta U a, one column each;
                           // column col_off[a] + j (a*attempts + j without col_off) is draw pos[chain(a)] + j
  const int* col_off;      // nullable [n_active + 1]: first pool column of every active chain; chain a may
                           // consume col_off[a+1] - col_off[a] momenta in this launch (instead of `attempts`)
  int* s_done;             // [n_chains] transitions completed
  int64_t* pos;            // [n_chains] momenta consumed (the chain's draw index)
  int* rcount;             // [n_chains] restarts inside the current transition
  int burn_in, n_per_chain;
  double* out;             // d x (n_chains * n_per_chain)
  uint64_t seed;
  int64_t chain_offset;    // global id of local chain c = chain_offset + c * chain_stride (Philox key)
  int chain_stride;        // 1, or F while the root chains of forked families run (see lqg_opts.fork)
  const double* injected;  // nullable [n_chains][n_injected]
  int64_t n_injected;
  int max_restarts;
  int* next_chain;         // work queue head (zeroed before launch)
  unsigned long long* stats;  // [0..7]: transitions,bounces,searches,vel,chk,evaluated,violations,failed;
                              // [11]: window passes without a hit (searched again 4 x wider);
                              // [10]: searches that fell back to tmg's exact hit-time arithmetic
  int* chain_status;       // [n_chains]
  int fast;                // 1: rank walls by tan(t/2) (hit_fast), exact arithmetic only near a threshold
  int window;              // 1: the fast search looks at a short time window first (see c_winK in lqg_hmc_box.cu)
};

cudaError_t launch_hmc_box(const BoxParams& p, int grid, cudaStream_t st);
cudaError_t launch_hit_time_batch(size_t n, const double* fa, const double* fb, const double* g, double klim,
                                  double* t_exact, int* cls, double* key, double* st, double* ct, cudaStream_t stream);
cudaError_t box_phase_cycles(unsigned long long* out16);   // diagnostic build (-DLQG_BOX_TIMING) only
int hmc_box_slots(int d);        // resident CTAs (= chains in flight) for dimension d, current device
cudaError_t launch_box_prep(int d, const double* mu, const double* lb, const double* ub,
                            const double* Sigma, double* glo, double* ghi, double* sdiag,
                            cudaStream_t st);
cudaError_t launch_box_init_state(int d, int n_chains, const double* init, int64_t init_ld,
                                  const double* mu, double* state, cudaStream_t st);
// forked families (lqg_opts.fork = F): chain c (global id g = g0 + c) takes the centred state of root
// (g - root0) / F; the first chain of a family (g % F == 0) continues the root's Philox stream at root_pos,
// the others start theirs at 0; every chain has s0 transitions behind it
cudaError_t launch_box_fork(int d, int n_chains, int F, int64_t g0, int64_t root0, const double* root_state,
                            const int64_t* root_pos, double* state, int64_t* pos, int* s_done, int s0, cudaStream_t st);
// n_active_out[0] = number of unfinished chains, n_active_out[1] = min over healthy chains of s_done.
// Then the next round is planned on the device: every unfinished chain gets as many pool columns as it
// can still use (transitions left + a restart allowance, at most the round's cap, which follows the
// slowest chain as before), col_off[0..n_active] = their exclusive prefix sums.
// host_mapped: device alias of 4 ints of mapped pinned host memory that receive
// {n_active, min s_done, pool columns of the next round, cap of the next round}
cudaError_t launch_box_compact(int n_chains, int total, const int* s_done, const int* status,
                               int* active_out, int* n_active_out, int* host_mapped, int max_attempts,
                               int short_tail, int* col_off, cudaStream_t st);
cudaError_t launch_box_violations(int d, size_t n_cols, const double* out, const double* lb,
                                  const double* ub, unsigned long long* count, cudaStream_t st);

// ---- structure-aware products of the simulate tail (lqg_sparse.cu) ------------------------------
// column c of a matrix whose columns come in blocks of w consecutive columns (stride ld) that start `pitch`
// doubles apart (a unit of chain-major samples); w = 0: plain columns with stride ld
struct ColBlocks {
  const double* base; int64_t ld; int w; int64_t pitch;
  __host__ __device__ const double* col(int64_t c) const {
    return w > 0 ? base + (size_t)(c / w) * pitch + (size_t)(c % w) * ld : base + (size_t)c * ld;
  }
};
// device-resident plan of xi = argmin |Lambda xi - eta| for a Lambda with short row windows
struct LsqPlanDev {
  int d, m;
  int W;                   // row window: non-zeros of row i lie in columns [rstart[i], rstart[i] + W)
  int Wc;                  // most non-zeros in a column
  int b;                   // half bandwidth of G = Lambda' Lambda
  const int* rstart;       // [d]
  const double* rval;      // [W x d]   rval[l*d + i] = Lambda[i, rstart[i] + l]
  const int* cidx;         // [Wc x m]  cidx[k*m + j] = k-th non-zero row of column j (-1: none)
  const double* cval;      // [Wc x m]
  const double* fw;        // [m x (b+1)] coefficients of the forward / backward recurrences of the banded
  const double* bw;        //             Cholesky factor of G (see lsq_band_chol_kernel)
  int refine;              // 1: one step of iterative refinement (cond(G) not provably small)
};
constexpr int kLsqMaxBand = 8, kLsqMaxColNnz = 32, kEllMaxNnz = 8;
// info[0] = widest row window - 1, info[1] = most non-zeros in a column (both atomicMax: zero them first)
cudaError_t launch_lsq_profile(int d, int m, const double* L, int* rstart, int* info, cudaStream_t st);
// fills the plan's arrays; info[2] = 1 + index of a non-positive pivot of G (0 = fine)
// gersh[0..1] = Gershgorin bounds of the spectrum of G
cudaError_t launch_lsq_build(int d, int m, int W, int Wc, int b, const double* L, const int* rstart, double* rval,
                             int* cidx, double* cval, double* lband, double* fw, double* bw, int* info, double* gersh,
                             cudaStream_t st);
cudaError_t launch_lsq_apply(const LsqPlanDev& P, ColBlocks eta, int64_t cols, double* xi, double* scratch, cudaStream_t st);
// first K non-zeros of every row of a dense slab (column order), info[0] = atomicMax of the row counts
cudaError_t launch_ell_from_dense(int nr, int m, int K, const double* P, int64_t ld, int* pidx, double* pval, int* info,
                                  cudaStream_t st);
cudaError_t launch_ell_spmm(int nr, int m, int K, const int* pidx, const double* pval, const double* X, int64_t ldx, int64_t N,
                            double* Y, int64_t ldy, cudaStream_t st);

// ---- canonical-frame HMC with dense constraints (lqg_hmc_canonical.cu) ------------------
struct CanonParams {
  int d, c;
  const double* F;         // c x d col-major (16-byte aligned, padded to an even element count)
  const double* g;         // [c]
  const double* frow2;     // [c] f.f per row
  const double* fnorm;     // [c] |f|
  double* state;           // d x n_chains (position b in the canonical frame), in/out
  int n_chains;
  int s_begin, s_end, burn_in, n_per_chain;
  double* out;             // d x (n_chains * n_per_chain)
  uint64_t seed;
  int64_t chain_offset;
  const double* injected;
  int64_t n_injected;
  int64_t* inj_pos;
  int max_restarts;
  int f_in_smem;           // stage F in shared memory with one bulk (TMA) copy per CTA
  // optional per-bounce trace (parity tooling): event e of chain ch -> trace_t/cn[ch*trace_cap + e]
  double* trace_t; int* trace_cn; int trace_cap; int* trace_n;   // trace_n[ch] = events seen
  int* next_chain;
  unsigned long long* stats;
  int* chain_status;
};
cudaError_t launch_hmc_canonical(const CanonParams& p, int grid, cudaStream_t st);
cudaError_t launch_canon_rownorms(int d, int c, const double* F, double* frow2, double* fnorm,
                                  cudaStream_t st);
cudaError_t launch_canon_violations(int d, int c, size_t n_cols, const double* F, const double* g,
                                    const double* fnorm, const double* out,
                                    unsigned long long* count, cudaStream_t st);
size_t canon_smem_bytes(int d, int c, int f_in_smem, int nch = 1);
int canon_f_fits_smem(int d, int c);

// ---- lock-step canonical HMC for large dense F (lqg_hmc_canonical_batched.cu) ----------------
struct CanonBatch {
  int d, c, n;                 // n chains
  const double* Ft;            // d x c: row k of F contiguous
  const double* g; const double* frow2; const double* fnorm;
  double* AB;                  // d x 2n: velocities a (columns 0..n-1), positions b (columns n..2n-1)
  double* Zlast;               // d x n: lastSample of every chain
  const double* FAB;           // c x 2n: F [A | B] of this round (GEMM output)
  double* tt; double* E;       // [n] remaining time, energy scale of the filter
  int* phase;                  // [n] 0 = transition starts, 2 = retry (both need a momentum), 1 = moving
  int* s_done; int* restart; int64_t* pos; int* status;
  int burn_in, n_per_chain, total;
  double* out;
  uint64_t seed; int64_t chain_offset;
  const double* injected; int64_t n_injected;
  int max_restarts;
  unsigned long long* stats;
  int* n_unfinished;           // chains that still have work after this round
};
cudaError_t launch_canon_batch_round_begin(const CanonBatch& p, cudaStream_t st);
cudaError_t launch_canon_batch_round_step(const CanonBatch& p, cudaStream_t st);

// ---- momentum pre-pass + generic FP64 tensor-core GEMM (lqg_dgemm.cu) -------------------
// Philox normals for the momentum pool: column a*attempts + j holds the d draws of
// (chain = active[a] (or a), draw index = pos[chain] + j)
struct NormalGen {
  uint64_t seed;
  int64_t chain_offset;
  int attempts;
  const int* active;       // nullable
  const int64_t* pos;      // [n_chains]
  const int* col_off;      // nullable [n_active + 1]: per-chain column ranges (see BoxParams::col_off)
  int n_active;
  int chain_stride;        // global id of local chain c = chain_offset + c * chain_stride
};
// b_blk_w > 0: the columns of B come in blocks of b_blk_w consecutive columns (stride ldb) that
// start b_blk_pitch doubles apart (a round of chain-major samples: w finished columns per chain)
cudaError_t launch_dgemm(int m, int n, int k, const double* A, int64_t lda, const double* B,
                         int64_t ldb, double* C, int64_t ldc, double* row_sum, int a_upper,
                         cudaStream_t st, int b_blk_w = 0, int64_t b_blk_pitch = 0);
cudaError_t launch_fill_normals(int d, size_t n_cols, double* A, const NormalGen& g, cudaStream_t st);
// Product summarised for the band statistics in the GEMM epilogue (rows = test points, columns = samples); see
// launch_paths_bands_fused (lqg_bands.cu), which owns the buffers
struct BandFuse {
  const unsigned long long* brackets = nullptr;   // [row][nprobs][2] order-preserving keys bracketing the wanted ranks
  unsigned* below = nullptr;                      // [row][nprobs] values below the bracket
  unsigned* cursor = nullptr;                     // [row][nprobs] values inside it (append cursor; may exceed cap)
  double* buf = nullptr;                          // [row][nprobs][cap] the values inside
  double* tile_sum = nullptr;                     // [dgemm_bands_parts(n)][rows] partial row sums, one per 32 columns
  unsigned cap = 0;
  int nprobs = 0;
};
bool dgemm_bands_usable(const double* A, int64_t lda, const double* B, int64_t ldb);
int dgemm_bands_parts(int n);
cudaError_t launch_dgemm_bands(int m, int n, int k, const double* A, int64_t lda, const double* B, int64_t ldb,
                               const BandFuse& f, cudaStream_t st);

// ---- dense set-up algebra (lqg_linalg.cu) ---------------------------------------------------
// C = alpha op(A) op(B) + beta C on the FP64 tensor cores; tri = 1: lower tiles only
cudaError_t launch_gemm(bool ta, bool tb, int m, int n, int k, double alpha, const double* A, int64_t lda,
                        const double* B, int64_t ldb, double beta, double* C, int64_t ldc, int tri,
                        cudaStream_t st);
// blocked Cholesky S = L L' (S: lower part read, destroyed); *info preset to a large value
cudaError_t launch_chol(int d, double* S, int64_t lds, double* L, int64_t ldl, int* info, cudaStream_t st);
// (L L') X = B in place (B d x nrhs overwritten with X)
cudaError_t launch_chol_solve(int d, const double* L, int64_t ldl, double* B, int64_t ldb, int nrhs,
                              cudaStream_t st);
cudaError_t launch_sym_reverse(int d, const double* in, int64_t ldi, double* out, int64_t ldo, int rev,
                               cudaStream_t st);
cudaError_t launch_reverse(int d, const double* in, double* out, cudaStream_t st);
cudaError_t launch_add_diag(int d, double* A, int64_t lda, double v, cudaStream_t st);
cudaError_t launch_transpose(int rows, int cols, const double* in, int64_t ldi, double* out, int64_t ldo,
                             cudaStream_t st);
cudaError_t launch_axpy(size_t n, double alpha, const double* x, double* y, cudaStream_t st);

// ---- ExpT set-up (lqg_expt_setup.cu): permuted Cholesky + Newton tilt, see the file header --------
int expt_setup_device(int d, double* Sigma, double* l, double* u, double* Lfull, double* L0, int* perm, double* y,
                      double* work, int* info_dev, double* psistar, int* iters, cudaStream_t st);

// ---- interior-point QP vector kernels (lqg_qp.cu); scal = 6 doubles -------------------------
cudaError_t launch_qp_init(int mc, const double* Ax, const double* b, double* s, double* lam, cudaStream_t st);
cudaError_t launch_qp_neg_copy(int n, const double* src, double* dst, cudaStream_t st);
cudaError_t launch_qp_resid(int n, int mc, const double* rd, const double* Ax, const double* s, const double* b,
                            const double* lam, double* rp, double* scal, cudaStream_t st);
cudaError_t launch_qp_scale_rows(int mc, int n, const double* A, const double* lam, const double* s, double* AW,
                                 cudaStream_t st);
cudaError_t launch_qp_prepare(int n, int mc, int corrector, double sigma_mu, const double* s, const double* lam,
                              const double* rp, const double* rd, double* ds, const double* dl, double* rc,
                              double* t, double* rhs, cudaStream_t st);
cudaError_t launch_qp_step(int mc, const double* s, const double* lam, const double* ds, const double* rc,
                           double* dl, double* scal, cudaStream_t st);
cudaError_t launch_qp_muaff(int mc, double a, const double* s, const double* lam, const double* ds,
                            const double* dl, double* scal, cudaStream_t st);
cudaError_t launch_qp_gather_rows(int n, int k, int mc, const double* A, const int* idx, double* AaT, cudaStream_t st);
cudaError_t launch_qp_update(int n, int mc, double a, double* x, const double* dx, double* s, const double* ds,
                             double* lam, const double* dl, cudaStream_t st);

// ---- hat basis (lqg_basis.cu) ------------------------------------------------------------------
cudaError_t launch_basis_hat(int n_t, int m, const double* x, const double* u, double* Phi, int64_t ld, int* flag,
                             cudaStream_t st);

// ---- RSM (lqg_rsm.cu) -------------------------------------------------------------------
struct RsmParams {
  int d;
  const double* map; const double* factor; const double* w; const double* lb; const double* ub;
  uint64_t seed;
  int64_t first_proposal;  // global index of the first proposal of this round
  int64_t n_proposals;     // proposals in this round
  const double* inj_z;     // nullable d x P
  const double* inj_u;     // nullable
  int64_t n_inj_u;
  int64_t inbox_before;    // in-box proposals in earlier rounds (rank offset of inj_u)
  unsigned char* flags;    // [n_proposals] bit0 = in box, bit1 = accepted
  double* tval;            // [n_proposals] exp((map - x)' w) for in-box proposals
};
int rsm_blocks(int64_t n_proposals);
// one round: evaluate, rank, accept, ordered compaction + emission
//   block_cnt [rsm_blocks], inbox_off / acc_off [rsm_blocks + 1] scratch; counters [2]
cudaError_t launch_rsm_round(const RsmParams& q, int* block_cnt, int64_t* inbox_off, int64_t* acc_off,
                             int64_t slot_base, int64_t nsim, double* out,
                             unsigned long long* counters, cudaStream_t st);

// ---- ExpT (lqg_expt.cu) ------------------------------------------------------------------
struct ExptParams {
  int d;
  const double* Lt;        // d x d, row k of the scaled zero-diagonal factor contiguous: Lt[j + k*d] = L0[k,j]
  const double* l; const double* u; const double* mu;   // scaled/permuted bounds, tilt (mu[d-1] = 0)
  double psistar;
  uint64_t seed;
  int64_t first_proposal, n_proposals;
  double* Z;               // coordinate-major proposals: Z[k*n_proposals + p]
  unsigned char* flags;    // [n_proposals] accepted
};
int expt_blocks(int64_t n);
cudaError_t launch_expt_transpose(int d, const double* L0, double* Lt, cudaStream_t st);
cudaError_t launch_expt_round(const ExptParams& q, int* block_cnt, int64_t* acc_off, int64_t slot_base,
                              int64_t nsim, double* out, unsigned long long* counters, cudaStream_t st);
cudaError_t launch_add_colvec(int m, size_t n, double* out, int64_t ld, const double* v, cudaStream_t st);

// ---- bands (lqg_bands.cu) ---------------------------------------------------------------
// ysim n_t x N col-major (ld).  scratch: N x n_t doubles (transposed copy).
// exact path (fallback): scratch = n_t * N doubles (transposed copy)
cudaError_t launch_bands_exact(int n_t, int64_t N, const double* ysim, int64_t ld, const double* probs_host,
                               int nprobs, double* mean, double* qtls, double* scratch, cudaStream_t st);
// bracketed path (one coalesced pass); scratch of bands_bracket_scratch_bytes() bytes (0: not applicable);
// *fail_dev -> n_t flags (inside scratch): 1 = this test point must go through launch_bands_exact
size_t bands_bracket_scratch_bytes(int n_t, int64_t N, int nprobs);
cudaError_t launch_bands_bracketed(int n_t, int64_t N, const double* ysim, int64_t ld, const double* probs_host,
                                   int nprobs, double* mean, double* qtls, void* scratch, int** fail_dev,
                                   cudaStream_t st);

// fused form: Phi_slab (n_t x m) X (m x N) is summarised inside the product's epilogue and never stored; brackets
// from a pilot product on the 2048 sampled columns.  scratch of paths_bands_fused_scratch_bytes() bytes
// (0: not applicable — few samples, operands the tensor-map kernel cannot take); *fail_dev as above
size_t paths_bands_fused_scratch_bytes(int n_t, int m, int64_t N, int nprobs, const double* phi, int64_t ldphi,
                                       const double* X, int64_t ldx);
cudaError_t launch_paths_bands_fused(int n_t, int m, int64_t N, const double* phi, int64_t ldphi, const double* X, int64_t ldx,
                                     const double* probs_host, int nprobs, double* mean, double* qtls, void* scratch,
                                     int** fail_dev, cudaStream_t st, cudaEvent_t ev_product_done);

// ---- peaks (lqg_peaks.cu) ---------------------------------------------------------------
cudaError_t measure_fp64_peaks(double* dfma_tflops, double* dmma_tflops, cudaStream_t st);
cudaError_t measure_fp64_mixed(double* tflops, double* dfma_share, cudaStream_t st);

}  // namespace lqg
