// Structure-aware forms of the two products of the simulate tail (R/lineqGP.R:639-640, R/lineqAGP.R:575):
//
//   xi.sim = qr.solve(Lambda, eta)      Lambda is what lineqGPR builds from its constraint blocks
//                                       (identity, first / second differences, block diagonal): every row
//                                       has its few non-zeros inside a short window.  Then G = Lambda' Lambda
//                                       is banded and the least-squares solution is a sparse product, a
//                                       banded Cholesky solve and one step of iterative refinement:
//                                       O(nnz + m b) per sample instead of the O(m d) of Lambda^+ eta.
//   ysim   = Phi.test %*% xi.sim        a 1-D hat basis has at most 2 non-zeros per row: an ELL product that
//                                       is bound by writing ysim, instead of an n_t x m x N GEMM.
//
// Both are HBM-bound kernels that leave the FP64 tensor pipe to the momentum GEMM running beside them.
// The callers (lqg_capi.cu) probe the structure on the device and fall back to the dense DMMA products
// (lqg_dgemm.cu) whenever it is not there (wide rows, dense Phi of an additive model).
#include <algorithm>

#include "lqg_common.cuh"
#include "lqg_kernels.h"

namespace lqg {

namespace {

// ---- structure probes ----------------------------------------------------------------------
// first and last non-zero column of every row; info[0] = max row width (last - first)
__global__ void lsq_row_profile_kernel(int d, int m, const double* __restrict__ L, int* __restrict__ rstart, int* info) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d) return;
  int lo = m, hi = -1;
  for (int j = 0; j < m; ++j)
    if (L[(size_t)j * d + i] != 0.0) { lo = min(lo, j); hi = max(hi, j); }
  rstart[i] = hi < 0 ? 0 : lo;
  atomicMax(info, hi < 0 ? 0 : hi - lo);
}
// non-zeros of every column (one warp per column); info[1] = max count
__global__ void lsq_col_profile_kernel(int d, int m, const double* __restrict__ L, int* info) {
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (j >= m) return;
  int cnt = 0;
  for (int i0 = 0; i0 < d; i0 += 32) {
    const int i = i0 + lane;
    cnt += __popc(__ballot_sync(0xffffffffu, i < d && L[(size_t)j * d + i] != 0.0));
  }
  if (lane == 0) atomicMax(info + 1, cnt);
}
// row windows: rval[l * d + i] = Lambda[i, rstart[i] + l] (0 beyond column m - 1)
__global__ void lsq_fill_rows_kernel(int d, int m, int W, const double* __restrict__ L, const int* __restrict__ rstart,
                                     double* __restrict__ rval) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d) return;
  const int s = rstart[i];
  for (int l = 0; l < W; ++l) rval[(size_t)l * d + i] = (s + l < m) ? L[(size_t)(s + l) * d + i] : 0.0;
}
// column lists in row order: cidx[k * m + j] = k-th non-zero row of column j (-1 = none), cval alike
__global__ void lsq_fill_cols_kernel(int d, int m, int Wc, const double* __restrict__ L, int* __restrict__ cidx,
                                     double* __restrict__ cval) {
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (j >= m) return;
  int base = 0;
  for (int i0 = 0; i0 < d; i0 += 32) {
    const int i = i0 + lane;
    const double v = i < d ? L[(size_t)j * d + i] : 0.0;
    const unsigned bal = __ballot_sync(0xffffffffu, v != 0.0);
    if (v != 0.0) {
      const int pos = base + __popc(bal & ((1u << lane) - 1u));
      if (pos < Wc) { cidx[(size_t)pos * m + j] = i; cval[(size_t)pos * m + j] = v; }
    }
    base += __popc(bal);
  }
  for (int k = base + lane; k < Wc; k += 32) { cidx[(size_t)k * m + j] = -1; cval[(size_t)k * m + j] = 0.0; }
}
// G = Lambda' Lambda, upper band: g[j * (b+1) + t] = G[j, j + t]
__global__ void lsq_gram_kernel(int d, int m, int Wc, int b, const double* __restrict__ L, const int* __restrict__ cidx,
                                const double* __restrict__ cval, double* __restrict__ g) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= m * (b + 1)) return;
  const int j = idx / (b + 1), t = idx - j * (b + 1);
  double acc = 0.0;
  if (j + t < m)
    for (int k = 0; k < Wc; ++k) {
      const int i = cidx[(size_t)k * m + j];
      if (i < 0) break;
      acc = fma(cval[(size_t)k * m + j], L[(size_t)(j + t) * d + i], acc);
    }
  g[idx] = acc;
}
// banded Cholesky G = L L' (one thread: m b^2 flops), written as the two coefficient tables of the solves:
//   forward   y_j = fw[j][0] r_j - sum_k fw[j][k] y_{j-k},   fw[j][0] = 1 / L[j,j],  fw[j][k] = L[j,j-k] / L[j,j]
//   backward  x_j = bw[j][0] y_j - sum_k bw[j][k] x_{j+k},   bw[j][0] = 1 / L[j,j],  bw[j][k] = L[j+k,j] / L[j,j]
// (tables indexed [j * (b+1) + k]; lb holds G's band on entry and is the work array).
// info[2] = 1 + index of a non-positive pivot (Lambda rank deficient).  gersh[0..1] = Gershgorin bounds of
// G's spectrum (min over rows of G_jj - sum |off-diagonal|, max of G_jj + sum): when they prove a small
// condition number the refinement step is not needed.
__global__ void lsq_band_chol_kernel(int m, int b, double* __restrict__ lb, double* __restrict__ fw, double* __restrict__ bw,
                                     int* info, double* gersh) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int ld = b + 1;
  double gmin = 1e300, gmax = 0.0;
  for (int j = 0; j < m; ++j) {
    double off = 0.0;
    for (int t = 1; t <= b; ++t) {
      if (j + t < m) off += fabs(lb[(size_t)j * ld + t]);                  // G[j, j+t]
      if (j - t >= 0) off += fabs(lb[(size_t)(j - t) * ld + t]);           // G[j-t, j]
    }
    gmin = fmin(gmin, lb[(size_t)j * ld] - off);
    gmax = fmax(gmax, lb[(size_t)j * ld] + off);
  }
  gersh[0] = gmin; gersh[1] = gmax;
  for (int j = 0; j < m; ++j) {
    // column j of L (lb[j * ld + t] = L[j + t, j]): subtract the contributions of columns j-k, k = 1..b
    for (int t = 0; t <= b && j + t < m; ++t) {
      double s = lb[(size_t)j * ld + t];
      for (int k = 1; k + t <= b && j - k >= 0; ++k) {
        const double ljk = lb[(size_t)(j - k) * ld + k];               // L[j, j-k]
        const double ltk = lb[(size_t)(j - k) * ld + k + t];           // L[j+t, j-k]
        s = fma(-ltk, ljk, s);
      }
      lb[(size_t)j * ld + t] = s;
    }
    const double piv = lb[(size_t)j * ld];
    if (!(piv > 0.0)) { if (info[2] == 0) info[2] = j + 1; lb[(size_t)j * ld] = 1.0; continue; }
    const double r = 1.0 / sqrt(piv);
    for (int t = 1; t <= b && j + t < m; ++t) lb[(size_t)j * ld + t] *= r;
    lb[(size_t)j * ld] = 1.0 / r;                                      // L[j, j]
  }
  for (int j = 0; j < m; ++j) {
    const double dinv = 1.0 / lb[(size_t)j * ld];
    fw[(size_t)j * ld] = dinv; bw[(size_t)j * ld] = dinv;
    for (int k = 1; k <= b; ++k) {
      fw[(size_t)j * ld + k] = j - k >= 0 ? lb[(size_t)(j - k) * ld + k] * dinv : 0.0;
      bw[(size_t)j * ld + k] = j + k < m ? lb[(size_t)j * ld + k] * dinv : 0.0;
    }
  }
}

// ---- per-sample kernels -----------------------------------------------------------------------
// out[:, c] = Lambda' v_c,  v_c = eta_c  or (xi given) the residual eta_c - Lambda xi_c
__global__ void lsq_rhs_kernel(int m, int d, int Wc, int W, const int* __restrict__ cidx, const double* __restrict__ cval,
                               const int* __restrict__ rstart, const double* __restrict__ rval, ColBlocks eta,
                               int64_t cols, const double* __restrict__ xi, double* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t c = blockIdx.y; c < cols; c += gridDim.y) {
    const double* e = eta.col(c);
    const double* x = xi ? xi + (size_t)c * m : nullptr;
    if (j < m) {
      double acc = 0.0;
      for (int k = 0; k < Wc; ++k) {
        const int i = cidx[(size_t)k * m + j];
        if (i < 0) break;
        double v = e[i];
        if (x) {
          const int rs = rstart[i];
          double s = 0.0;
          for (int l = 0; l < W; ++l) s = fma(rval[(size_t)l * d + i], x[min(rs + l, m - 1)], s);
          v -= s;
        }
        acc = fma(cval[(size_t)k * m + j], v, acc);
      }
      out[(size_t)c * m + j] = acc;
    }
  }
}

// X[:, c] <- (L L')^{-1} X[:, c] for 32 columns per CTA.  The recurrences run along a column (warp 0: one
// thread per column), the memory is contiguous along a column too: 32 x 32 tiles go through shared memory so
// that both the global accesses (tile rows) and the recurrences (tile columns) are conflict free.  Three tile
// buffers: while warp 0 solves tile k, warps 1-3 fetch tile k+1 and write tile k-1 back, one barrier per tile.
// The term of the most recent unknown comes last in every recurrence: its dependent chain is one FMA long.
template <int B>
__global__ void __launch_bounds__(128) lsq_band_solve_kernel(int m, const double* __restrict__ fw, const double* __restrict__ bw,
                                                             int64_t cols, double* __restrict__ X) {
  __shared__ double tile[3][32][33];
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
  const int64_t c0 = (int64_t)blockIdx.x * 32;
  constexpr int LD = B + 1;
  const int nt = (m + 31) / 32;
  auto load = [&](int k, int j0) {            // warps 1..3
    for (int cc = wy - 1; cc < 32; cc += 3)
      if (c0 + cc < cols && j0 + lane < m) tile[k % 3][cc][lane] = X[(size_t)(c0 + cc) * m + j0 + lane];
  };
  auto store = [&](int k, int j0) {
    for (int cc = wy - 1; cc < 32; cc += 3)
      if (c0 + cc < cols && j0 + lane < m) X[(size_t)(c0 + cc) * m + j0 + lane] = tile[k % 3][cc][lane];
  };
  double prev[B];
  for (int pass = 0; pass < 2; ++pass) {      // 0: forward L y = r (tiles ascending), 1: backward L' x = y (descending)
    const double* __restrict__ cf = pass == 0 ? fw : bw;
    auto j0_of = [&](int k) { return pass == 0 ? k * 32 : (nt - 1 - k) * 32; };
#pragma unroll
    for (int k = 0; k < B; ++k) prev[k] = 0.0;
    if (wy > 0) load(0, j0_of(0));
    __syncthreads();
    for (int k = 0; k < nt; ++k) {
      if (wy > 0) {
        if (k + 1 < nt) load(k + 1, j0_of(k + 1));
        if (k > 0) store(k - 1, j0_of(k - 1));
      } else {
        const int j0 = j0_of(k);
        const int tn = min(32, m - j0);
        double* col = tile[k % 3][lane];
        for (int q = 0; q < tn; ++q) {
          const int t = pass == 0 ? q : tn - 1 - q;
          const double* c = cf + (size_t)(j0 + t) * LD;
          double s = __ldg(c) * col[t];
#pragma unroll
          for (int kk = B; kk >= 1; --kk) s = fma(-__ldg(c + kk), prev[kk - 1], s);
#pragma unroll
          for (int kk = B - 1; kk > 0; --kk) prev[kk] = prev[kk - 1];
          prev[0] = s;
          col[t] = s;
        }
      }
      __syncthreads();
    }
    if (wy > 0) store(nt - 1, j0_of(nt - 1));
    __syncthreads();                          // the backward pass reads what the forward pass stored (same threads
                                              // write and read an element only within one CTA: global memory is coherent per CTA after the barrier)
  }
}

__global__ void axpy1_kernel(size_t n, const double* __restrict__ x, double* __restrict__ y) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) y[i] += x[i];
}

// ---- ELL form of a slab of Phi and its product ------------------------------------------------------
// pidx[k * nr + t] / pval: the first K non-zeros of row t, in column order; info[0] = max non-zeros of a row
__global__ void ell_from_dense_kernel(int nr, int m, int K, const double* __restrict__ P, int64_t ld, int* __restrict__ pidx,
                                      double* __restrict__ pval, int* info) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nr) return;
  int cnt = 0;
  for (int j = 0; j < m; ++j) {
    const double v = P[(size_t)j * ld + t];
    if (v != 0.0) {
      if (cnt < K) { pidx[(size_t)cnt * nr + t] = j; pval[(size_t)cnt * nr + t] = v; }
      ++cnt;
    }
  }
  for (int k = cnt; k < K; ++k) { pidx[(size_t)k * nr + t] = 0; pval[(size_t)k * nr + t] = 0.0; }
  atomicMax(info, cnt);
}
// Y[t, s] = sum_k pval[k, t] X[pidx[k, t], s]
template <int K>
__global__ void __launch_bounds__(256) ell_spmm_kernel(int nr, int m, const int* __restrict__ pidx, const double* __restrict__ pval,
                                                       const double* __restrict__ X, int64_t ldx, int64_t N, double* __restrict__ Y, int64_t ldy) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  int idx[K]; double val[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    idx[k] = t < nr ? pidx[(size_t)k * nr + t] : 0;
    val[k] = t < nr ? pval[(size_t)k * nr + t] : 0.0;
  }
  for (int64_t s = blockIdx.y; s < N; s += gridDim.y) {
    const double* x = X + (size_t)s * ldx;
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) acc = fma(val[k], __ldg(x + idx[k]), acc);
    if (t < nr) __stcs(Y + (size_t)s * ldy + t, acc);
  }
}

}  // namespace

cudaError_t launch_lsq_profile(int d, int m, const double* L, int* rstart, int* info, cudaStream_t st) {
  lsq_row_profile_kernel<<<(d + 127) / 128, 128, 0, st>>>(d, m, L, rstart, info);
  lsq_col_profile_kernel<<<(m * 32 + 255) / 256, 256, 0, st>>>(d, m, L, info);
  launch_counter() += 2;
  return cudaGetLastError();
}

cudaError_t launch_lsq_build(int d, int m, int W, int Wc, int b, const double* L, const int* rstart, double* rval,
                             int* cidx, double* cval, double* lband, double* fw, double* bw, int* info, double* gersh,
                             cudaStream_t st) {
  lsq_fill_rows_kernel<<<(d + 127) / 128, 128, 0, st>>>(d, m, W, L, rstart, rval);
  lsq_fill_cols_kernel<<<(m * 32 + 255) / 256, 256, 0, st>>>(d, m, Wc, L, cidx, cval);
  lsq_gram_kernel<<<(m * (b + 1) + 255) / 256, 256, 0, st>>>(d, m, Wc, b, L, cidx, cval, lband);
  lsq_band_chol_kernel<<<1, 32, 0, st>>>(m, b, lband, fw, bw, info, gersh);
  launch_counter() += 4;
  return cudaGetLastError();
}

static cudaError_t band_solve(int m, int b, const double* fw, const double* bw, int64_t cols, double* X, cudaStream_t st) {
  const unsigned grid = (unsigned)((cols + 31) / 32);
  switch (b) {
    case 1: lsq_band_solve_kernel<1><<<grid, 128, 0, st>>>(m, fw, bw, cols, X); break;
    case 2: lsq_band_solve_kernel<2><<<grid, 128, 0, st>>>(m, fw, bw, cols, X); break;
    case 3: lsq_band_solve_kernel<3><<<grid, 128, 0, st>>>(m, fw, bw, cols, X); break;
    case 4: lsq_band_solve_kernel<4><<<grid, 128, 0, st>>>(m, fw, bw, cols, X); break;
    case 5: lsq_band_solve_kernel<5><<<grid, 128, 0, st>>>(m, fw, bw, cols, X); break;
    case 6: lsq_band_solve_kernel<6><<<grid, 128, 0, st>>>(m, fw, bw, cols, X); break;
    case 7: lsq_band_solve_kernel<7><<<grid, 128, 0, st>>>(m, fw, bw, cols, X); break;
    case 8: lsq_band_solve_kernel<8><<<grid, 128, 0, st>>>(m, fw, bw, cols, X); break;
    default: return cudaErrorInvalidValue;
  }
  ++launch_counter();
  return cudaGetLastError();
}

// xi[:, c] = argmin |Lambda xi - eta_c| for c < cols: normal equations by the banded factor and, unless
// Gershgorin's bounds prove cond(G) small (P.refine = 0), one step of iterative refinement with the true
// residual (then it matches the QR solution of qr.solve to rounding for the condition numbers lineqGPR's
// Lambda has).  scratch: m x cols doubles (refinement only).
cudaError_t launch_lsq_apply(const LsqPlanDev& P, ColBlocks eta, int64_t cols, double* xi, double* scratch, cudaStream_t st) {
  if (cols <= 0) return cudaSuccess;
  const dim3 grid((P.m + 127) / 128, (unsigned)std::min<int64_t>(cols, 148 * 16));
  lsq_rhs_kernel<<<grid, 128, 0, st>>>(P.m, P.d, P.Wc, P.W, P.cidx, P.cval, P.rstart, P.rval, eta, cols, nullptr, xi);
  ++launch_counter();
  cudaError_t e = band_solve(P.m, P.b, P.fw, P.bw, cols, xi, st);
  if (e != cudaSuccess || !P.refine) return e;
  lsq_rhs_kernel<<<grid, 128, 0, st>>>(P.m, P.d, P.Wc, P.W, P.cidx, P.cval, P.rstart, P.rval, eta, cols, xi, scratch);
  ++launch_counter();
  e = band_solve(P.m, P.b, P.fw, P.bw, cols, scratch, st);
  if (e != cudaSuccess) return e;
  axpy1_kernel<<<148 * 8, 256, 0, st>>>((size_t)P.m * cols, scratch, xi);
  ++launch_counter();
  return cudaGetLastError();
}

cudaError_t launch_ell_from_dense(int nr, int m, int K, const double* P, int64_t ld, int* pidx, double* pval, int* info,
                                  cudaStream_t st) {
  ell_from_dense_kernel<<<(nr + 127) / 128, 128, 0, st>>>(nr, m, K, P, ld, pidx, pval, info);
  ++launch_counter();
  return cudaGetLastError();
}

cudaError_t launch_ell_spmm(int nr, int m, int K, const int* pidx, const double* pval, const double* X, int64_t ldx, int64_t N,
                            double* Y, int64_t ldy, cudaStream_t st) {
  if (nr <= 0 || N <= 0) return cudaSuccess;
  const dim3 grid((nr + 255) / 256, (unsigned)std::min<int64_t>(N, 32768));
  if (K <= 2) ell_spmm_kernel<2><<<grid, 256, 0, st>>>(nr, m, pidx, pval, X, ldx, N, Y, ldy);
  else if (K <= 4) ell_spmm_kernel<4><<<grid, 256, 0, st>>>(nr, m, pidx, pval, X, ldx, N, Y, ldy);
  else if (K <= 8) ell_spmm_kernel<8><<<grid, 256, 0, st>>>(nr, m, pidx, pval, X, ldx, N, Y, ldy);
  else return cudaErrorInvalidValue;
  ++launch_counter();
  return cudaGetLastError();
}

}  // namespace lqg
