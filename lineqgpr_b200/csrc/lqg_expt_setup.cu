// Set-up of the minimax exponential-tilting sampler on the device — what TruncatedNormal::mvrandn does
// before its accept-reject loop (tmvrnorm.ExpT, R/lineqGPsamplers.R:179-185; Botev 2017, JRSS-B 79(1),
// Algorithm 1 and eq. 11), the same arithmetic as the host restatement lineqgpr_b200/expt.py:
//
//   cholperm   Cholesky factor with the variable reordering of Gibson, Glasbey & Elston: at step j the
//              remaining variable with the smallest conditional truncated mass comes next.  Inherently
//              sequential in j; one CTA runs the whole factorization with one row per thread: the
//              running sums  ss_i = sum_k L_ik^2  and  cc_i = sum_k L_ik z_k  are kept per row, so a
//              step costs O(d) for the pivot choice (log-mass of every remaining row, block arg-min
//              with the first index winning ties) and one dot product of length j per row for the new
//              column (L column-major: coalesced across rows; row j broadcast from shared memory).
//   nleq       Newton iteration for grad psi(x, mu) = 0.  The Jacobian [[A, B'], [B, D]] has a diagonal
//              D = diag(1 + dP) > 0 and A = L0' diag(dP) L0 <= 0, so the step comes from the SPD Schur
//              complement  M = -A + B' D^-1 B  (two DMMA GEMMs, blocked Cholesky, lqg_linalg.cu):
//                  M dx = g1 - B' D^-1 g2 ,   dmu = D^-1 (-g2 - B dx).
#include <math.h>

#include <algorithm>
#include <vector>

#include "lqg_common.cuh"
#include "lqg_kernels.h"

namespace lqg {

namespace {

constexpr double kEpsD = 2.220446049250313e-16;
constexpr double kInvSqrt2Pi = 0.3989422804014327;
constexpr int CP_T = 1024;

// ---- permuted Cholesky ----------------------------------------------------------------------------
// S: d x d (destroyed: rows/columns are swapped in place), L: d x d output (lower), l/u in-out (permuted),
// perm out, z scratch [d]; info[0] = 1 if Sigma is not positive semi-definite (pivot < -0.001)
__global__ void __launch_bounds__(CP_T) expt_cholperm_kernel(int d, double* __restrict__ S, double* __restrict__ L,
                                                              double* __restrict__ l, double* __restrict__ u,
                                                              int* __restrict__ perm, double* __restrict__ ss,
                                                              double* __restrict__ cc, int* info) {
  extern __shared__ double rowj[];                 // [d] row j of L (columns < j)
  __shared__ double red_v[CP_T / 32];
  __shared__ int red_i[CP_T / 32];
  __shared__ int k_sm;
  __shared__ double zj_sm, ljj_sm;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < d; i += CP_T) { perm[i] = i; ss[i] = 0.0; cc[i] = 0.0; }
  for (size_t e = tid; e < (size_t)d * d; e += CP_T) L[e] = 0.0;
  __syncthreads();
  for (int j = 0; j < d; ++j) {
    // (1) conditional log-mass of every remaining variable, arg-min (first index wins ties)
    double bv = INFINITY; int bi = 0x7fffffff;
    for (int i = j + tid; i < d; i += CP_T) {
      double s = S[(size_t)i * d + i] - ss[i];
      s = sqrt(s < 0.0 ? kEpsD : s);
      const double c = cc[i];
      const double pr = lnNpr((l[i] - c) / s, (u[i] - c) / s);
      if (pr < bv || (pr == bv && i < bi)) { bv = pr; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { red_v[warp] = bv; red_i[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
      double v = red_v[0]; int k = red_i[0];
      for (int w = 1; w < CP_T / 32; ++w)
        if (red_v[w] < v || (red_v[w] == v && red_i[w] < k)) { v = red_v[w]; k = red_i[w]; }
      k_sm = (k == 0x7fffffff) ? j : k;            // all NaN: keep the order (the pivot test below reports)
    }
    __syncthreads();
    const int k = k_sm;
    // (2) swap variables j <-> k everywhere: rows then columns of S, rows of L, bounds, bookkeeping
    if (k != j) {
      for (int c = tid; c < d; c += CP_T) {
        const double a = S[(size_t)c * d + j], b = S[(size_t)c * d + k];
        S[(size_t)c * d + j] = b; S[(size_t)c * d + k] = a;
      }
      __syncthreads();
      for (int r = tid; r < d; r += CP_T) {
        const double a = S[(size_t)j * d + r], b = S[(size_t)k * d + r];
        S[(size_t)j * d + r] = b; S[(size_t)k * d + r] = a;
      }
      for (int c = tid; c < j; c += CP_T) {
        const double a = L[(size_t)c * d + j], b = L[(size_t)c * d + k];
        L[(size_t)c * d + j] = b; L[(size_t)c * d + k] = a;
      }
      if (tid == 0) {
        double t; int ti;
        t = l[j]; l[j] = l[k]; l[k] = t;
        t = u[j]; u[j] = u[k]; u[k] = t;
        t = ss[j]; ss[j] = ss[k]; ss[k] = t;
        t = cc[j]; cc[j] = cc[k]; cc[k] = t;
        ti = perm[j]; perm[j] = perm[k]; perm[k] = ti;
      }
      __syncthreads();
    }
    // (3) pivot, row j of L into shared memory
    for (int c = tid; c < j; c += CP_T) rowj[c] = L[(size_t)c * d + j];
    if (tid == 0) {
      double s = S[(size_t)j * d + j] - ss[j];
      if (s < -0.001) info[0] = 1;
      if (!(s >= 0.0)) s = kEpsD;
      ljj_sm = sqrt(s);
    }
    __syncthreads();
    const double ljj = ljj_sm;
    if (tid == 0) L[(size_t)j * d + j] = ljj;
    // (4) column j: L[i, j] = (S[i, j] - L[i, :j] . L[j, :j]) / L[j, j]
    for (int i = j + 1 + tid; i < d; i += CP_T) {
      double a0 = 0.0, a1 = 0.0;
      int c = 0;
      for (; c + 1 < j; c += 2) {
        a0 = fma(L[(size_t)c * d + i], rowj[c], a0);
        a1 = fma(L[(size_t)(c + 1) * d + i], rowj[c + 1], a1);
      }
      if (c < j) a0 = fma(L[(size_t)c * d + i], rowj[c], a0);
      const double v = (S[(size_t)j * d + i] - (a0 + a1)) / ljj;
      L[(size_t)j * d + i] = v;
      ss[i] = fma(v, v, ss[i]);
    }
    // (5) the expectation of the truncated variable z_j, then cc_i += L[i, j] z_j
    if (tid == 0) {
      const double c = cc[j];
      const double tl = (l[j] - c) / ljj, tu = (u[j] - c) / ljj;
      const double w = lnNpr(tl, tu);
      zj_sm = (exp(-0.5 * tl * tl - w) - exp(-0.5 * tu * tu - w)) * kInvSqrt2Pi;
    }
    __syncthreads();
    const double zj = zj_sm;
    for (int i = j + 1 + tid; i < d; i += CP_T) cc[i] = fma(L[(size_t)j * d + i], zj, cc[i]);
    __syncthreads();
  }
}

// L0 = L / diag(L) - I (row scaling, zero diagonal), l, u scaled alike, dg = diag(L)
__global__ void expt_scale_kernel(int d, const double* __restrict__ L, double* __restrict__ L0, double* __restrict__ l,
                                  double* __restrict__ u) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)d * d) return;
  const int i = (int)(idx % d), c = (int)(idx / d);
  const double dg = L[(size_t)i * d + i];
  L0[idx] = (c == i) ? 0.0 : L[idx] / dg;
  if (c == 0) { l[i] /= dg; u[i] /= dg; }
}

// ---- Newton iteration for the tilt ---------------------------------------------------------------------
// y = [x (n) | mu (n)], n = d - 1.  One pass over the rows (one thread each): c = L0 x, the truncated-normal
// terms, the gradient pieces that do not need a column sum, and the scaled matrices W = diag(sqrt(-dP)) L0
// (d x n) and V = diag(Dg^-1/2) B (n x n), B = (-I + diag(dP) L0)[:n, :n].
__global__ void expt_rows_kernel(int d, const double* __restrict__ L0, const double* __restrict__ l, const double* __restrict__ u,
                                 const double* __restrict__ y, double* __restrict__ P, double* __restrict__ dP,
                                 double* __restrict__ W, double* __restrict__ V, double* __restrict__ psi_terms) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d) return;
  const int n = d - 1;
  double c = 0.0;
  for (int k = 0; k < i && k < n; ++k) c = fma(L0[(size_t)k * d + i], y[k], c);     // L0 is lower triangular, zero diagonal
  const double xi = i < n ? y[i] : 0.0, mui = i < n ? y[n + i] : 0.0;
  double lt = l[i] - mui - c, ut = u[i] - mui - c;
  const double w = lnNpr(lt, ut);
  const double pl = exp(-0.5 * lt * lt - w) * kInvSqrt2Pi, pu = exp(-0.5 * ut * ut - w) * kInvSqrt2Pi;
  const double Pi = pl - pu;
  if (isinf(lt)) lt = 0.0;
  if (isinf(ut)) ut = 0.0;
  const double dPi = -Pi * Pi + lt * pl - ut * pu;
  P[i] = Pi; dP[i] = dPi;
  psi_terms[i] = w + 0.5 * mui * mui - xi * mui;
  const double sw = sqrt(fmax(-dPi, 0.0));
  const double sv = i < n ? 1.0 / sqrt(1.0 + dPi) : 0.0;
  for (int k = 0; k < n; ++k) {
    const double lv = L0[(size_t)k * d + i];
    W[(size_t)k * d + i] = sw * lv;
    if (i < n) V[(size_t)k * n + i] = sv * ((k == i ? -1.0 : 0.0) + dPi * lv);
  }
}
// g1 = -mu + L0[:, :n]' P,  g2 = mu - x + P[:n];  rhs1 = g1 - B' Dg^-1 g2 is finished by expt_rhs_kernel
__global__ void expt_grad_kernel(int d, const double* __restrict__ L0, const double* __restrict__ y, const double* __restrict__ P,
                                 const double* __restrict__ dP, double* __restrict__ g, double* __restrict__ t2) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = d - 1;
  if (k >= n) return;
  double acc = 0.0;
  for (int i = k + 1; i < d; ++i) acc = fma(P[i], L0[(size_t)k * d + i], acc);   // column k of L0 below the diagonal
  g[k] = -y[n + k] + acc;
  const double g2 = y[n + k] - y[k] + P[k];
  g[n + k] = g2;
  t2[k] = g2 / (1.0 + dP[k]);                                                      // Dg^-1 g2
}
// rhs[k] = g1[k] - (B' t2)[k],  B = (-I + diag(dP) L0)[:n, :n]  ->  (B' t2)[k] = -t2[k] + sum_i dP[i] L0[i, k] t2[i]
__global__ void expt_rhs_kernel(int d, const double* __restrict__ L0, const double* __restrict__ dP, const double* __restrict__ g,
                                const double* __restrict__ t2, double* __restrict__ rhs) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = d - 1;
  if (k >= n) return;
  double acc = -t2[k];
  for (int i = k + 1; i < n; ++i) acc = fma(dP[i] * L0[(size_t)k * d + i], t2[i], acc);
  rhs[k] = g[k] - acc;
}
// dmu = Dg^-1 (-g2 - B dx);  y += [dx | dmu]
__global__ void expt_step_kernel(int d, const double* __restrict__ L0, const double* __restrict__ dP, const double* __restrict__ g,
                                 const double* __restrict__ dx, double* __restrict__ y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = d - 1;
  if (i >= n) return;
  double bdx = -dx[i];
  for (int k = 0; k < i; ++k) bdx = fma(dP[i] * L0[(size_t)k * d + i], dx[k], bdx);
  const double dmu = (-g[n + i] - bdx) / (1.0 + dP[i]);
  y[i] += dx[i];
  y[n + i] += dmu;
}
__global__ void expt_negate_kernel(int n, double* x) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = -x[i];
}

}  // namespace

// Everything on `st`; all pointers device memory.  Sigma is destroyed.  Outputs: Lfull (d x d), L0 (d x d),
// l / u (scaled, permuted, in place), perm [d], y = [x | mu] (2 (d-1)), *psistar, *iters.
// work: doubles, at least 3 d^2 + 16 d.   Returns 0, 1 (Sigma not PSD), 2 (Newton did not converge),
// 3 (Schur complement not positive definite) or a negative cudaError.
int expt_setup_device(int d, double* Sigma, double* l, double* u, double* Lfull, double* L0, int* perm, double* y,
                      double* work, int* info_dev, double* psistar, int* iters, cudaStream_t st) {
  const int n = d - 1;
  double* ss = work; double* cc = ss + d;
  double* P = cc + d; double* dP = P + d; double* psi_t = dP + d; double* g = psi_t + d; double* t2 = g + 2 * d;
  double* rhs = t2 + d;
  double* W = rhs + 2 * d + (8 - (size_t)(10 * d) % 8) % 8;      // keep the matrices 64-byte aligned
  double* V = W + (size_t)d * d; double* M = V + (size_t)d * d;
  cudaError_t e;
#define EX_TRY(x) do { e = (x); if (e != cudaSuccess) return -(int)e; } while (0)
  EX_TRY(cudaMemsetAsync(info_dev, 0, 2 * sizeof(int), st));
  EX_TRY(cudaFuncSetAttribute(expt_cholperm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(d * sizeof(double))));
  expt_cholperm_kernel<<<1, CP_T, d * sizeof(double), st>>>(d, Sigma, Lfull, l, u, perm, ss, cc, info_dev);
  expt_scale_kernel<<<(unsigned)(((size_t)d * d + 255) / 256), 256, 0, st>>>(d, Lfull, L0, l, u);
  launch_counter() += 2;
  EX_TRY(cudaGetLastError());
  int h_info[2] = {0, 0};
  EX_TRY(cudaMemcpyAsync(h_info, info_dev, sizeof(h_info), cudaMemcpyDeviceToHost, st));
  EX_TRY(cudaStreamSynchronize(st));
  if (h_info[0]) return 1;
  EX_TRY(cudaMemsetAsync(y, 0, (size_t)2 * std::max(n, 1) * sizeof(double), st));
  *iters = 0;
  std::vector<double> h_g((size_t)2 * std::max(n, 1)), h_psi(d);
  bool converged = n == 0;
  for (int it = 0; it < 101 && !converged; ++it) {
    expt_rows_kernel<<<(d + 63) / 64, 64, 0, st>>>(d, L0, l, u, y, P, dP, W, V, psi_t);
    expt_grad_kernel<<<(n + 63) / 64, 64, 0, st>>>(d, L0, y, P, dP, g, t2);
    launch_counter() += 2;
    EX_TRY(cudaMemcpyAsync(h_g.data(), g, (size_t)2 * n * sizeof(double), cudaMemcpyDeviceToHost, st));
    EX_TRY(cudaStreamSynchronize(st));
    double gg = 0.0;
    for (double v : h_g) gg += v * v;
    ++*iters;
    // expt.nleq: the step is taken first, then the norm of the gradient it started from is tested
    // M = W[:, :n]' W[:, :n] + V' V  (lower tiles), SPD
    EX_TRY(launch_gemm(true, false, n, n, d, 1.0, W, d, W, d, 0.0, M, n, 1, st));
    EX_TRY(launch_gemm(true, false, n, n, n, 1.0, V, n, V, n, 1.0, M, n, 1, st));
    expt_rhs_kernel<<<(n + 63) / 64, 64, 0, st>>>(d, L0, dP, g, t2, rhs);
    ++launch_counter();
    EX_TRY(cudaMemsetAsync(info_dev + 1, 0x7f, sizeof(int), st));
    EX_TRY(launch_chol(n, M, n, W, n, info_dev + 1, st));                 // W is free again: receives the factor
    EX_TRY(launch_chol_solve(n, W, n, rhs, n, 1, st));
    // Newton: y += Jac^-1 (-grad); with M = -(A - B' D^-1 B):  M dx = g1 - B' D^-1 g2  gives dx directly
    expt_step_kernel<<<(n + 63) / 64, 64, 0, st>>>(d, L0, dP, g, rhs, y);
    ++launch_counter();
    EX_TRY(cudaMemcpyAsync(h_info, info_dev, sizeof(h_info), cudaMemcpyDeviceToHost, st));
    EX_TRY(cudaStreamSynchronize(st));
    if (h_info[1] != 0x7f7f7f7f) return 3;
    if (gg <= 1e-10) converged = true;
  }
  if (!converged) return 2;
  // psi* at the solution
  expt_rows_kernel<<<(d + 63) / 64, 64, 0, st>>>(d, L0, l, u, y, P, dP, W, V, psi_t);
  ++launch_counter();
  EX_TRY(cudaMemcpyAsync(h_psi.data(), psi_t, (size_t)d * sizeof(double), cudaMemcpyDeviceToHost, st));
  EX_TRY(cudaStreamSynchronize(st));
  double ps = 0.0;
  for (double v : h_psi) ps += v;
  *psistar = ps;
#undef EX_TRY
  return 0;
}

}  // namespace lqg
