// Hat-basis matrix Phi: basisCompute.lineqGP, R/lineqGP.R:29-76, and the
// per-dimension blocks of PhiAll.test, R/lineqAGP.R:341-348.
//   Phi[t, j] = max(1 - |x_t - u_j| / h, 0),  h = u_{j+1} - u_j to the right of the knot
//                                                (x_t - u_j > 1e-12, :49-50),
//                                             h = u_j - u_{j-1} to the left (x_t - u_j < 0, :51-52)
// evaluated with the same IEEE operations as the R expression (one division, one subtraction),
// so the result is bit-identical to the host restatement.  R indexes u out of range for a point
// outside [u_1, u_m] and fails on the NA; here the flag is raised and the call returns an error.
// HBM-bound: 8 bytes written per element, x and u stay in L1/L2.
#include "lqg_common.cuh"
#include "lqg_kernels.h"

namespace lqg {

__global__ void basis_hat_kernel(int n_t, int m, const double* __restrict__ x, const double* __restrict__ u,
                                 double* __restrict__ Phi, int64_t ld, int* flag) {
  const size_t total = (size_t)n_t * m;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int t = (int)(idx % n_t), j = (int)(idx / n_t);
    const double diff = x[t] - u[j];
    double dist = 0.0;
    bool bad = diff != diff;
    if (diff > 1e-12) {
      if (j + 1 < m) dist = fabs(diff / (u[j + 1] - u[j])); else bad = true;
    } else if (diff < 0.0) {
      if (j > 0) dist = fabs(diff / (u[j] - u[j - 1])); else bad = true;
    }
    if (bad) { *flag = 1; dist = 2.0; }
    Phi[(size_t)j * ld + t] = dist <= 1.0 ? 1.0 - dist : 0.0;
  }
}

cudaError_t launch_basis_hat(int n_t, int m, const double* x, const double* u, double* Phi, int64_t ld, int* flag,
                             cudaStream_t st) {
  const size_t total = (size_t)n_t * m;
  if (total == 0) return cudaSuccess;
  const unsigned grid = (unsigned)min((size_t)148 * 16, (total + 255) / 256);
  basis_hat_kernel<<<grid, 256, 0, st>>>(n_t, m, x, u, Phi, ld, flag);
  ++launch_counter();
  return cudaGetLastError();
}

}  // namespace lqg
