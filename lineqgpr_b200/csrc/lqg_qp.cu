// Vector kernels of the constrained-MAP quadratic programme
//   min 1/2 x'Dx - dvec'x   s.t.  A x >= b
// (quadprog::solve.QP(Dmat, dvec, t(A), b) at R/lineqGP.R:530 and R/lineqAGP.R:431), solved on
// the device by a Mehrotra predictor-corrector interior-point method.  The O(n^2 c) and O(n^3)
// work per iteration — H = D + A' diag(lam/s) A, its Cholesky factorisation and the two solves —
// runs through lqg_linalg.cu (FP64 tensor cores); what is here are the O(n + c) vector updates
// and the deterministic single-CTA reductions between them.  The iteration itself is driven
// from the host (lqg_solve_qp in lqg_capi.cu) with three small read-backs per iteration.
#include "lqg_common.cuh"
#include "lqg_kernels.h"

namespace lqg {

namespace {

constexpr int RT = 1024;

__device__ __forceinline__ double block_reduce(double v, int op /*0 sum, 1 max, 2 min*/) {
  __shared__ double red[32];
  __syncthreads();
  auto comb = [op](double a, double b) { return op == 0 ? a + b : (op == 1 ? fmax(a, b) : fmin(a, b)); };
  for (int o = 16; o > 0; o >>= 1) v = comb(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    double w = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : (op == 0 ? 0.0 : (op == 1 ? -INFINITY : INFINITY));
    for (int o = 16; o > 0; o >>= 1) w = comb(w, __shfl_xor_sync(0xffffffffu, w, o));
    if (threadIdx.x == 0) red[0] = w;
  }
  __syncthreads();
  return red[0];
}

// s = max(Ax - b, 1), lam = 1
__global__ void qp_init_kernel(int mc, const double* Ax, const double* b, double* s, double* lam) {
  for (int i = threadIdx.x; i < mc; i += RT) { s[i] = fmax(Ax[i] - b[i], 1.0); lam[i] = 1.0; }
}
// rd = -dvec (the products D x and -A' lam are accumulated onto it by GEMMs)
__global__ void qp_neg_copy_kernel(int n, const double* src, double* dst) {
  for (int i = threadIdx.x; i < n; i += RT) dst[i] = -src[i];
}
// rp = Ax - s - b;  scal[0] = max|rd|, scal[1] = max|rp|, scal[2] = s.lam
__global__ void qp_resid_kernel(int n, int mc, const double* rd, const double* Ax, const double* s, const double* b,
                                const double* lam, double* rp, double* scal) {
  double m_rd = 0.0, m_rp = 0.0, dot = 0.0;
  for (int i = threadIdx.x; i < n; i += RT) m_rd = fmax(m_rd, fabs(rd[i]));
  for (int i = threadIdx.x; i < mc; i += RT) {
    const double r = Ax[i] - s[i] - b[i];
    rp[i] = r;
    m_rp = fmax(m_rp, fabs(r));
    dot = fma(s[i], lam[i], dot);
  }
  m_rd = block_reduce(m_rd, 1);
  m_rp = block_reduce(m_rp, 1);
  dot = block_reduce(dot, 0);
  if (threadIdx.x == 0) { scal[0] = m_rd; scal[1] = m_rp; scal[2] = dot; }
}
// AW[i, j] = A[i, j] lam[i] / s[i]
__global__ void qp_scale_rows_kernel(int mc, int n, const double* A, const double* lam, const double* s, double* AW) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)mc * n) return;
  const int i = (int)(idx % mc);
  AW[idx] = A[idx] * (lam[i] / s[i]);
}
// complementarity residual and right-hand side pieces of one Newton step:
//   rc = s lam (+ ds dl - sigma_mu for the corrector, ds/dl still holding the affine step)
//   t = (-rc - lam rp)/s;  rhs = -rd;  ds = rp   (A' t and A dx are then accumulated by GEMMs)
__global__ void qp_prepare_kernel(int n, int mc, int corrector, double sigma_mu, const double* s, const double* lam,
                                  const double* rp, const double* rd, double* ds, const double* dl, double* rc,
                                  double* t, double* rhs) {
  for (int i = threadIdx.x; i < mc; i += RT) {
    double r = s[i] * lam[i];
    if (corrector) r = r + ds[i] * dl[i] - sigma_mu;
    rc[i] = r;
    t[i] = (-r - lam[i] * rp[i]) / s[i];
    ds[i] = rp[i];
  }
  for (int i = threadIdx.x; i < n; i += RT) rhs[i] = -rd[i];
}
// dl = (-rc - lam ds)/s;  scal[3] = min over ds<0 of -s/ds, scal[4] = min over dl<0 of -lam/dl (Inf if none)
__global__ void qp_step_kernel(int mc, const double* s, const double* lam, const double* ds, const double* rc,
                               double* dl, double* scal) {
  double a_s = INFINITY, a_l = INFINITY;
  for (int i = threadIdx.x; i < mc; i += RT) {
    const double dsi = ds[i];
    const double dli = (-rc[i] - lam[i] * dsi) / s[i];
    dl[i] = dli;
    if (dsi < 0.0) a_s = fmin(a_s, -s[i] / dsi);
    if (dli < 0.0) a_l = fmin(a_l, -lam[i] / dli);
  }
  a_s = block_reduce(a_s, 2);
  a_l = block_reduce(a_l, 2);
  if (threadIdx.x == 0) { scal[3] = a_s; scal[4] = a_l; }
}
// scal[5] = sum (s + a ds)(lam + a dl)
__global__ void qp_muaff_kernel(int mc, double a, const double* s, const double* lam, const double* ds,
                                const double* dl, double* scal) {
  double acc = 0.0;
  for (int i = threadIdx.x; i < mc; i += RT) acc = fma(s[i] + a * ds[i], lam[i] + a * dl[i], acc);
  acc = block_reduce(acc, 0);
  if (threadIdx.x == 0) scal[5] = acc;
}
__global__ void qp_update_kernel(int n, int mc, double a, double* x, const double* dx, double* s, const double* ds,
                                 double* lam, const double* dl) {
  for (int i = threadIdx.x; i < n; i += RT) x[i] = fma(a, dx[i], x[i]);
  for (int i = threadIdx.x; i < mc; i += RT) { s[i] = fma(a, ds[i], s[i]); lam[i] = fma(a, dl[i], lam[i]); }
}

// AaT[j, q] = A[idx[q], j]: the active rows of A (mc x n), transposed into an n x k matrix
__global__ void qp_gather_rows_kernel(int n, int k, int mc, const double* __restrict__ A, const int* __restrict__ idx,
                                      double* __restrict__ AaT) {
  const size_t total = (size_t)n * k;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(e % n), q = (int)(e / n);
    AaT[e] = A[(size_t)j * mc + idx[q]];
  }
}

}  // namespace

cudaError_t launch_qp_gather_rows(int n, int k, int mc, const double* A, const int* idx, double* AaT, cudaStream_t st) {
  const size_t total = (size_t)n * k;
  if (total == 0) return cudaSuccess;
  const unsigned grid = (unsigned)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  qp_gather_rows_kernel<<<grid, 256, 0, st>>>(n, k, mc, A, idx, AaT);
  ++launch_counter();
  return cudaGetLastError();
}

#define LQG_QP_LAUNCH(call) do { call; ++launch_counter(); return cudaGetLastError(); } while (0)

cudaError_t launch_qp_init(int mc, const double* Ax, const double* b, double* s, double* lam, cudaStream_t st) {
  LQG_QP_LAUNCH((qp_init_kernel<<<1, RT, 0, st>>>(mc, Ax, b, s, lam)));
}
cudaError_t launch_qp_neg_copy(int n, const double* src, double* dst, cudaStream_t st) {
  LQG_QP_LAUNCH((qp_neg_copy_kernel<<<1, RT, 0, st>>>(n, src, dst)));
}
cudaError_t launch_qp_resid(int n, int mc, const double* rd, const double* Ax, const double* s, const double* b,
                            const double* lam, double* rp, double* scal, cudaStream_t st) {
  LQG_QP_LAUNCH((qp_resid_kernel<<<1, RT, 0, st>>>(n, mc, rd, Ax, s, b, lam, rp, scal)));
}
cudaError_t launch_qp_scale_rows(int mc, int n, const double* A, const double* lam, const double* s, double* AW,
                                 cudaStream_t st) {
  const size_t total = (size_t)mc * n;
  LQG_QP_LAUNCH((qp_scale_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(mc, n, A, lam, s, AW)));
}
cudaError_t launch_qp_prepare(int n, int mc, int corrector, double sigma_mu, const double* s, const double* lam,
                              const double* rp, const double* rd, double* ds, const double* dl, double* rc,
                              double* t, double* rhs, cudaStream_t st) {
  LQG_QP_LAUNCH((qp_prepare_kernel<<<1, RT, 0, st>>>(n, mc, corrector, sigma_mu, s, lam, rp, rd, ds, dl, rc, t, rhs)));
}
cudaError_t launch_qp_step(int mc, const double* s, const double* lam, const double* ds, const double* rc,
                           double* dl, double* scal, cudaStream_t st) {
  LQG_QP_LAUNCH((qp_step_kernel<<<1, RT, 0, st>>>(mc, s, lam, ds, rc, dl, scal)));
}
cudaError_t launch_qp_muaff(int mc, double a, const double* s, const double* lam, const double* ds,
                            const double* dl, double* scal, cudaStream_t st) {
  LQG_QP_LAUNCH((qp_muaff_kernel<<<1, RT, 0, st>>>(mc, a, s, lam, ds, dl, scal)));
}
cudaError_t launch_qp_update(int n, int mc, double a, double* x, const double* dx, double* s, const double* ds,
                             double* lam, const double* dl, cudaStream_t st) {
  LQG_QP_LAUNCH((qp_update_kernel<<<1, RT, 0, st>>>(n, mc, a, x, dx, s, ds, lam, dl)));
}

}  // namespace lqg
