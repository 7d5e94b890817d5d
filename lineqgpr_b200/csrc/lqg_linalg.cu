// Dense FP64 set-up algebra on the device: everything the reference
// does on the host before and after the sampler call, as blocked algorithms whose flops all
// run through one FP64 tensor-core GEMM (mma.sync.m8n8k4.f64, SASS DMMA):
//   Sigma_eta = Lambda Sigma Lambda' + nugget I            R/lineqGP.R:611-614, R/lineqAGP.R:558-566
//   R = chol(M), Ri = solve(R)  ->  U with U U' = Sigma    tmg:R/tmg.R:15-28 (see whiten_box)
//   positive-definiteness test                             tmg:R/tmg.R:17-20, R/lineqGP.R:527-528
//   qr.solve(Lambda, eta)  ->  Lambda^+                    R/lineqGP.R:639, R/lineqAGP.R:575
//   solve(A, B) for symmetric positive definite A          R/lineqGP.R:514-523 (posterior algebra)
//
// Blocked right-looking Cholesky, block size 64:
//   diagonal block   one CTA factors S_kk = L_kk L_kk' in shared memory
//   panel            L_ik L_kk' = S_ik by forward substitution, one warp per panel row
//   trailing update  S_ij -= L_ik L_jk'  (lower tiles only) (GEMM, alpha = -1, beta = 1)
// The SPD solve A X = B is two blocked substitutions: a warp-per-column triangular solve with
// the 64 x 64 diagonal block, then a GEMM update of the remaining rows.  The triangular solves
// are real substitutions, not products with an inverted diagonal block: the matrices of this
// path (Sigma_eta with a 1e-7 nugget, the interior-point KKT matrix near convergence) have
// condition numbers of 1e10 and more, which leaves two or three digits of margin on the pivots;
// an explicit block inverse is not backward stable there.
#include "lqg_common.cuh"
#include "lqg_kernels.h"

namespace lqg {

namespace {

constexpr int GM = 64, GN = 64, GK = 16, GT = 128;
constexpr int GA_LD = GM + 4;    // As[kk * GA_LD + r]
constexpr int GB_LD = GK + 4;    // Bs[n * GB_LD + kk]
constexpr int NB = 64;           // Cholesky block size (= GN: a panel is one tile column)

__device__ __forceinline__ void dmma884g(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// C (m x n) = alpha op(A) op(B) + beta C, column-major; op(A) is m x k, op(B) is k x n.
// tri = 1: only tiles that touch the lower triangle (col0 <= row0 + GM - 1) are computed.
template <bool TA, bool TB>
__global__ void __launch_bounds__(GT)
gemm_general_kernel(int m, int n, int k, double alpha, const double* __restrict__ A, int64_t lda,
                    const double* __restrict__ B, int64_t ldb, double beta, double* C, int64_t ldc,
                    int tri) {
  __shared__ double As[2][GK * GA_LD];
  __shared__ double Bs[2][GN * GB_LD];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 1, wn = warp >> 1;
  const int row0 = blockIdx.x * GM, col0 = blockIdx.y * GN;
  if (tri == 1 && col0 > row0 + GM - 1) return;

  double ra[8], rb[8];
  const int kt_end = (k + GK - 1) / GK;
  auto load_tile = [&](int kt) {
    const int k0 = kt * GK;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int e = tid + q * GT;
      int r, kk;
      if (TA) { kk = e % GK; r = e / GK; } else { r = e % GM; kk = e / GM; }
      const int gr = row0 + r, gk = k0 + kk;
      ra[q] = (gr < m && gk < k) ? __ldg(TA ? A + (size_t)gr * lda + gk : A + (size_t)gk * lda + gr) : 0.0;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int e = tid + q * GT;
      int nn, kk;
      if (TB) { nn = e % GN; kk = e / GN; } else { kk = e % GK; nn = e / GK; }
      const int gc = col0 + nn, gk = k0 + kk;
      rb[q] = (gc < n && gk < k) ? __ldg(TB ? B + (size_t)gk * ldb + gc : B + (size_t)gc * ldb + gk) : 0.0;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int e = tid + q * GT;
      int r, kk;
      if (TA) { kk = e % GK; r = e / GK; } else { r = e % GM; kk = e / GM; }
      As[buf][kk * GA_LD + r] = ra[q];
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int e = tid + q * GT;
      int nn, kk;
      if (TB) { nn = e % GN; kk = e / GN; } else { kk = e % GK; nn = e / GK; }
      Bs[buf][nn * GB_LD + kk] = rb[q];
    }
  };

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  if (kt_end > 0) {
    load_tile(0);
    store_tile(0);
  }
  __syncthreads();
  const int fr = lane >> 2, fk = lane & 3;
  int buf = 0;
  for (int kt = 0; kt < kt_end; ++kt) {
    if (kt + 1 < kt_end) load_tile(kt + 1);
    const double* as = As[buf];
    const double* bs = Bs[buf];
#pragma unroll
    for (int k4 = 0; k4 < GK / 4; ++k4) {
      double af[4], bf[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) af[i] = as[(k4 * 4 + fk) * GA_LD + wm * 32 + i * 8 + fr];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = bs[(wn * 32 + j * 8 + fr) * GB_LD + k4 * 4 + fk];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884g(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
    if (kt + 1 < kt_end) store_tile(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = row0 + wm * 32 + i * 8 + fr;
    if (r >= m) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int cc = col0 + wn * 32 + j * 8 + 2 * fk + h;
        if (cc < n) {
          double* dst = C + (size_t)cc * ldc + r;
          const double v = alpha * acc[i][j][h];
          *dst = (beta == 0.0) ? v : fma(beta, *dst, v);
        }
      }
    }
  }
}

// One CTA: S_kk (nbk x nbk, lower part read) = L L'; writes L into Lout (lower part, zeros above).
// A non-positive or NaN pivot records its 1-based global index in *info (smallest wins) and is
// replaced by 1.
__global__ void __launch_bounds__(256) chol_diag_kernel(int nbk, int k0, const double* S, int64_t lds,
                                                        double* Lout, int64_t ldl, int* info) {
  __shared__ double s[NB * 65];      // s[i * 65 + l]
  const int tid = threadIdx.x;
  for (int e = tid; e < NB * NB; e += 256) {
    const int i = e % NB, l = e / NB;
    double v = (i == l) ? 1.0 : 0.0;
    if (i < nbk && l < nbk && l <= i) v = S[(size_t)l * lds + i];
    s[i * 65 + l] = v;
  }
  __syncthreads();
  // thread -> (row i, column phase lq): column l = j+1+lq, j+5+lq, ... of row i; a warp shares l
  // (broadcast read of s[l][j]) and walks 32 consecutive rows (stride 65: conflict-free)
  const int ri = tid & (NB - 1), lq = tid >> 6;
  for (int j = 0; j < nbk; ++j) {
    if (tid == 0) {
      double piv = s[j * 65 + j];
      if (!(piv > 0.0) || isinf(piv)) { atomicMin(info, k0 + j + 1); piv = 1.0; }
      s[j * 65 + j] = sqrt(piv);
    }
    __syncthreads();
    const double ljj = s[j * 65 + j];
    if (tid > j && tid < nbk) s[tid * 65 + j] /= ljj;
    __syncthreads();
    if (ri > j && ri < nbk) {
      const double lij = s[ri * 65 + j];
      for (int l = j + 1 + lq; l <= ri; l += 4) s[ri * 65 + l] = fma(-lij, s[l * 65 + j], s[ri * 65 + l]);
    }
    __syncthreads();
  }
  for (int e = tid; e < nbk * nbk; e += 256) {
    const int i = e % nbk, l = e / nbk;
    Lout[(size_t)l * ldl + i] = (l <= i) ? s[i * 65 + l] : 0.0;
  }
}

// Triangular solves with one 64 x 64 diagonal block, one warp per right-hand side:
//   trans = 0:  L x = b   (forward substitution)      trans = 1:  L' x = b   (backward)
// Right-hand side r has its element j at B[j * sj + r * sr]; the solution is written to
// X[j * xj + r * xr] (may alias B).  Lane l owns elements l and l + 32; the pivot element is
// scaled by the reciprocal diagonal (computed once per CTA) in its owner lane and broadcast.
__global__ void __launch_bounds__(256) tri_solve_kernel(int nbk, const double* __restrict__ Lkk, int64_t ldl, int trans,
                                                        int nrhs, const double* B, int64_t sj, int64_t sr,
                                                        double* X, int64_t xj, int64_t xr) {
  __shared__ double s[NB * 65];      // s[i * 65 + l] = L[i, l]
  __shared__ double rdiag[NB];       // 1 / L[j, j]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int e = tid; e < NB * NB; e += 256) {
    const int i = e % NB, l = e / NB;
    s[i * 65 + l] = (i < nbk && l <= i) ? Lkk[(size_t)l * ldl + i] : (i == l ? 1.0 : 0.0);
  }
  __syncthreads();
  if (tid < NB) rdiag[tid] = 1.0 / s[tid * 65 + tid];
  __syncthreads();
  for (int r = blockIdx.x * 8 + warp; r < nrhs; r += gridDim.x * 8) {
    double b0 = lane < nbk ? B[(size_t)lane * sj + (size_t)r * sr] : 0.0;
    double b1 = lane + 32 < nbk ? B[(size_t)(lane + 32) * sj + (size_t)r * sr] : 0.0;
    const int i0 = lane, i1 = lane + 32;
    if (!trans) {
      for (int j = 0; j < nbk; ++j) {
        const double mine = (j < 32 ? b0 : b1) * rdiag[j];
        const double xv = __shfl_sync(0xffffffffu, mine, j & 31);
        if (i0 == j) b0 = xv; else if (i0 > j) b0 = fma(-s[i0 * 65 + j], xv, b0);
        if (i1 == j) b1 = xv; else if (i1 > j) b1 = fma(-s[i1 * 65 + j], xv, b1);
      }
    } else {
      for (int j = nbk - 1; j >= 0; --j) {                    // L'[i, j] = L[j, i]
        const double mine = (j < 32 ? b0 : b1) * rdiag[j];
        const double xv = __shfl_sync(0xffffffffu, mine, j & 31);
        if (i0 == j) b0 = xv; else if (i0 < j) b0 = fma(-s[j * 65 + i0], xv, b0);
        if (i1 == j) b1 = xv; else if (i1 < j) b1 = fma(-s[j * 65 + i1], xv, b1);
      }
    }
    if (lane < nbk) X[(size_t)lane * xj + (size_t)r * xr] = b0;
    if (lane + 32 < nbk) X[(size_t)(lane + 32) * xj + (size_t)r * xr] = b1;
  }
}

// out = (in + in')/2 with both index ranges reversed when rev (J S J), d x d
__global__ void sym_reverse_kernel(int d, const double* in, int64_t ldi, double* out, int64_t ldo, int rev) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)d * d) return;
  const int i = (int)(idx % d), j = (int)(idx / d);
  const int si = rev ? d - 1 - i : i, sj = rev ? d - 1 - j : j;
  out[(size_t)j * ldo + i] = 0.5 * (in[(size_t)sj * ldi + si] + in[(size_t)si * ldi + sj]);
}
// out[i, j] = in[d-1-i, d-1-j]
__global__ void reverse_kernel(int d, const double* in, double* out) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)d * d) return;
  const int i = (int)(idx % d), j = (int)(idx / d);
  out[idx] = in[(size_t)(d - 1 - j) * d + (d - 1 - i)];
}
__global__ void add_diag_kernel(int d, double* A, int64_t lda, double v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < d) A[(size_t)i * lda + i] += v;
}
// out (cols x rows) = in' (in is rows x cols)
__global__ void transpose_kernel(int rows, int cols, const double* in, int64_t ldi, double* out, int64_t ldo) {
  __shared__ double t[32][33];
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int q = threadIdx.y; q < 32; q += 8) {
    const int r = r0 + threadIdx.x, c = c0 + q;
    t[q][threadIdx.x] = (r < rows && c < cols) ? in[(size_t)c * ldi + r] : 0.0;
  }
  __syncthreads();
  for (int q = threadIdx.y; q < 32; q += 8) {
    const int c = c0 + threadIdx.x, r = r0 + q;
    if (r < rows && c < cols) out[(size_t)r * ldo + c] = t[threadIdx.x][q];
  }
}

__global__ void axpy_kernel(size_t n, double alpha, const double* x, double* y) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = fma(alpha, x[i], y[i]);
}

}  // namespace

cudaError_t launch_axpy(size_t n, double alpha, const double* x, double* y, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  axpy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, alpha, x, y);
  ++launch_counter();
  return cudaGetLastError();
}

cudaError_t launch_gemm(bool ta, bool tb, int m, int n, int k, double alpha, const double* A, int64_t lda,
                        const double* B, int64_t ldb, double beta, double* C, int64_t ldc, int tri,
                        cudaStream_t st) {
  if (m <= 0 || n <= 0) return cudaSuccess;
  dim3 grid((m + GM - 1) / GM, (n + GN - 1) / GN);
  if (grid.y > 65535u) return cudaErrorInvalidValue;
  if (!ta && !tb) gemm_general_kernel<false, false><<<grid, GT, 0, st>>>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, tri);
  else if (ta && !tb) gemm_general_kernel<true, false><<<grid, GT, 0, st>>>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, tri);
  else if (!ta && tb) gemm_general_kernel<false, true><<<grid, GT, 0, st>>>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, tri);
  else gemm_general_kernel<true, true><<<grid, GT, 0, st>>>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, tri);
  ++launch_counter();
  return cudaGetLastError();
}

static cudaError_t launch_tri_solve(int nbk, const double* Lkk, int64_t ldl, int trans, int nrhs, const double* B,
                                    int64_t sj, int64_t sr, double* X, int64_t xj, int64_t xr, cudaStream_t st) {
  if (nrhs <= 0) return cudaSuccess;
  const int grid = (nrhs + 7) / 8 < 148 * 4 ? (nrhs + 7) / 8 : 148 * 4;
  tri_solve_kernel<<<grid, 256, 0, st>>>(nbk, Lkk, ldl, trans, nrhs, B, sj, sr, X, xj, xr);
  ++launch_counter();
  return cudaGetLastError();
}

// S (d x d, lower part, destroyed) -> L (d x d, lower, strict upper part zeroed).
// *info (device int, preset to a large value) receives the first non-positive pivot (1-based) if any.
cudaError_t launch_chol(int d, double* S, int64_t lds, double* L, int64_t ldl, int* info, cudaStream_t st) {
  cudaError_t e = cudaMemset2DAsync(L, ldl * sizeof(double), 0, (size_t)d * sizeof(double), d, st);
  if (e != cudaSuccess) return e;
  for (int k0 = 0; k0 < d; k0 += NB) {
    const int nbk = d - k0 < NB ? d - k0 : NB;
    const int k1 = k0 + nbk;
    double* Lkk = L + (size_t)k0 * ldl + k0;
    chol_diag_kernel<<<1, 256, 0, st>>>(nbk, k0, S + (size_t)k0 * lds + k0, lds, Lkk, ldl, info);
    ++launch_counter();
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    const int below = d - k1;
    if (below > 0) {
      // panel: L[k1:, k0:k1] L_kk' = S[k1:, k0:k1]; row r of the panel is right-hand side r
      e = launch_tri_solve(nbk, Lkk, ldl, 0, below, S + (size_t)k0 * lds + k1, lds, 1,
                           L + (size_t)k0 * ldl + k1, ldl, 1, st);
      if (e != cudaSuccess) return e;
      // trailing: S[k1:, k1:] -= P P'   (lower tiles)
      const double* P = L + (size_t)k0 * ldl + k1;
      e = launch_gemm(false, true, below, below, nbk, -1.0, P, ldl, P, ldl, 1.0,
                      S + (size_t)k1 * lds + k1, lds, 1, st);
      if (e != cudaSuccess) return e;
    }
  }
  return cudaSuccess;
}

// Solves (L L') X = B in place for nrhs right-hand sides (B d x nrhs, overwritten with X).
cudaError_t launch_chol_solve(int d, const double* L, int64_t ldl, double* B, int64_t ldb, int nrhs,
                              cudaStream_t st) {
  cudaError_t e;
  for (int k0 = 0; k0 < d; k0 += NB) {                       // L Y = B
    const int nbk = d - k0 < NB ? d - k0 : NB, k1 = k0 + nbk;
    e = launch_tri_solve(nbk, L + (size_t)k0 * ldl + k0, ldl, 0, nrhs, B + k0, 1, ldb, B + k0, 1, ldb, st);
    if (e != cudaSuccess) return e;
    if (d - k1 > 0) {
      e = launch_gemm(false, false, d - k1, nrhs, nbk, -1.0, L + (size_t)k0 * ldl + k1, ldl, B + k0, ldb, 1.0,
                      B + k1, ldb, 0, st);
      if (e != cudaSuccess) return e;
    }
  }
  const int last = ((d - 1) / NB) * NB;
  for (int k0 = last; k0 >= 0; k0 -= NB) {                   // L' X = Y
    const int nbk = d - k0 < NB ? d - k0 : NB;
    e = launch_tri_solve(nbk, L + (size_t)k0 * ldl + k0, ldl, 1, nrhs, B + k0, 1, ldb, B + k0, 1, ldb, st);
    if (e != cudaSuccess) return e;
    if (k0 > 0) {
      e = launch_gemm(true, false, k0, nrhs, nbk, -1.0, L + k0, ldl, B + k0, ldb, 1.0, B, ldb, 0, st);
      if (e != cudaSuccess) return e;
    }
  }
  return cudaSuccess;
}

cudaError_t launch_sym_reverse(int d, const double* in, int64_t ldi, double* out, int64_t ldo, int rev,
                               cudaStream_t st) {
  const size_t total = (size_t)d * d;
  sym_reverse_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d, in, ldi, out, ldo, rev);
  ++launch_counter();
  return cudaGetLastError();
}
cudaError_t launch_reverse(int d, const double* in, double* out, cudaStream_t st) {
  const size_t total = (size_t)d * d;
  reverse_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d, in, out);
  ++launch_counter();
  return cudaGetLastError();
}
cudaError_t launch_add_diag(int d, double* A, int64_t lda, double v, cudaStream_t st) {
  add_diag_kernel<<<(d + 255) / 256, 256, 0, st>>>(d, A, lda, v);
  ++launch_counter();
  return cudaGetLastError();
}
cudaError_t launch_transpose(int rows, int cols, const double* in, int64_t ldi, double* out, int64_t ldo,
                             cudaStream_t st) {
  dim3 grid((rows + 31) / 32, (cols + 31) / 32), block(32, 8);
  transpose_kernel<<<grid, block, 0, st>>>(rows, cols, in, ldi, out, ldo);
  ++launch_counter();
  return cudaGetLastError();
}

}  // namespace lqg
