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
//   diagonal block   one CTA factors S_kk = L_kk L_kk' in shared memory and also forms
//                    W_k = L_kk^-1, so that every triangular solve below becomes a GEMM
//   panel            L_ik = S_ik W_k'                       (GEMM, op(B) = transpose)
//   trailing update  S_ij -= L_ik L_jk'  (lower tiles only) (GEMM, alpha = -1, beta = 1)
// The SPD solve A X = B is two blocked substitutions with the same W_k blocks.
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

// One CTA: S_kk (nbk x nbk, lower part read) = L L'; writes L into Lout (lower part) and
// W = L^-1 (lower, 64 x 64 column-major, identity-padded) into Wout.  A non-positive or NaN
// pivot records its 1-based global index in *info (smallest wins) and is replaced by 1.
__global__ void __launch_bounds__(256) chol_diag_kernel(int nbk, int k0, const double* S, int64_t lds,
                                                        double* Lout, int64_t ldl, double* Wout, int* info) {
  extern __shared__ double cd_sm[];
  double* s = cd_sm;                 // s[i * 65 + l]
  double* w = cd_sm + NB * 65;       // w[l * 64 + c] = W[l, c]
  const int tid = threadIdx.x;
  for (int e = tid; e < NB * NB; e += 256) {
    const int i = e % NB, l = e / NB;
    double v = (i == l) ? 1.0 : 0.0;
    if (i < nbk && l < nbk && l <= i) v = S[(size_t)l * lds + i];
    s[i * 65 + l] = v;
  }
  __syncthreads();
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
    // trailing update of the lower triangle: (i, l) with j < l <= i < nbk
    const int rem = nbk - j - 1;
    for (int e = tid; e < rem * rem; e += 256) {
      const int i = j + 1 + e % rem, l = j + 1 + e / rem;
      if (l <= i) s[i * 65 + l] = fma(-s[i * 65 + j], s[l * 65 + j], s[i * 65 + l]);
    }
    __syncthreads();
  }
  for (int e = tid; e < nbk * nbk; e += 256) {
    const int i = e % nbk, l = e / nbk;
    Lout[(size_t)l * ldl + i] = (l <= i) ? s[i * 65 + l] : 0.0;
  }
  // W = L^-1 by forward substitution, one column per thread (uniform loops: broadcast reads of L)
  if (tid < NB) {
    const int c = tid;
    for (int i = 0; i < NB; ++i) {
      double acc = (i == c) ? 1.0 : 0.0;
      for (int l = 0; l < i; ++l) acc = fma(-s[i * 65 + l], w[l * NB + c], acc);
      w[i * NB + c] = acc / s[i * 65 + i];
    }
  }
  __syncthreads();
  for (int e = tid; e < NB * NB; e += 256) {
    const int i = e % NB, c = e / NB;
    Wout[(size_t)c * NB + i] = (c <= i) ? w[i * NB + c] : 0.0;
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

size_t chol_winv_elems(int d) { return (size_t)((d + NB - 1) / NB) * NB * NB; }

// S (d x d, lower part, destroyed) -> L (d x d, lower, strict upper part zeroed), Winv blocks.
// *info (device int, preset to INT_MAX) receives the first non-positive pivot (1-based) if any.
cudaError_t launch_chol(int d, double* S, int64_t lds, double* L, int64_t ldl, double* Winv, int* info,
                        cudaStream_t st) {
  static const cudaError_t attr = cudaFuncSetAttribute(chol_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                       (int)((NB * 65 + NB * NB) * sizeof(double)));
  if (attr != cudaSuccess) return attr;
  cudaError_t e = cudaMemset2DAsync(L, ldl * sizeof(double), 0, (size_t)d * sizeof(double), d, st);
  if (e != cudaSuccess) return e;
  for (int k0 = 0; k0 < d; k0 += NB) {
    const int nbk = d - k0 < NB ? d - k0 : NB;
    const int k1 = k0 + nbk;
    double* W = Winv + (size_t)(k0 / NB) * NB * NB;
    chol_diag_kernel<<<1, 256, (NB * 65 + NB * NB) * sizeof(double), st>>>(
        nbk, k0, S + (size_t)k0 * lds + k0, lds, L + (size_t)k0 * ldl + k0, ldl, W, info);
    ++launch_counter();
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    const int below = d - k1;
    if (below > 0) {
      // panel: L[k1:, k0:k1] = S[k1:, k0:k1] W'
      e = launch_gemm(false, true, below, nbk, nbk, 1.0, S + (size_t)k0 * lds + k1, lds, W, NB, 0.0,
                      L + (size_t)k0 * ldl + k1, ldl, 0, st);
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

// Solves (L L') X = B for nrhs right-hand sides.  B (d x nrhs) is overwritten with X; T is a
// d x nrhs scratch.  Forward: T_k = W_k B_k, B_{k+1:} -= L_{k+1:,k} T_k.  Backward:
// B_k = W_k' T_k, T_{:k} -= L_{k,:k}' B_k.
cudaError_t launch_chol_solve(int d, const double* L, int64_t ldl, const double* Winv, double* B, int64_t ldb,
                              double* T, int64_t ldt, int nrhs, cudaStream_t st) {
  cudaError_t e;
  for (int k0 = 0; k0 < d; k0 += NB) {
    const int nbk = d - k0 < NB ? d - k0 : NB, k1 = k0 + nbk;
    const double* W = Winv + (size_t)(k0 / NB) * NB * NB;
    e = launch_gemm(false, false, nbk, nrhs, nbk, 1.0, W, NB, B + k0, ldb, 0.0, T + k0, ldt, 0, st);
    if (e != cudaSuccess) return e;
    if (d - k1 > 0) {
      e = launch_gemm(false, false, d - k1, nrhs, nbk, -1.0, L + (size_t)k0 * ldl + k1, ldl, T + k0, ldt, 1.0,
                      B + k1, ldb, 0, st);
      if (e != cudaSuccess) return e;
    }
  }
  const int last = ((d - 1) / NB) * NB;
  for (int k0 = last; k0 >= 0; k0 -= NB) {
    const int nbk = d - k0 < NB ? d - k0 : NB;
    const double* W = Winv + (size_t)(k0 / NB) * NB * NB;
    e = launch_gemm(true, false, nbk, nrhs, nbk, 1.0, W, NB, T + k0, ldt, 0.0, B + k0, ldb, 0, st);
    if (e != cudaSuccess) return e;
    if (k0 > 0) {
      e = launch_gemm(true, false, k0, nrhs, nbk, -1.0, L + k0, ldl, B + k0, ldb, 1.0, T, ldt, 0, st);
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
