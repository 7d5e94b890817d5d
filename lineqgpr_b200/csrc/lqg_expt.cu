// Minimax exponential-tilting sampler: the per-sample part of TruncatedNormal::mvrandn behind
//   tmvrnorm.ExpT     R/lineqGPsamplers.R:179-185
// (TruncatedNormal is not vendored in the reference tree; this follows the published algorithm,
//  Botev (2017) JRSS-B 79(1):125-148, Algorithm 1 and its sequential proposal / accept step.)
//
// Host (lineqgpr_b200/expt.py, as in the R package): permuted Cholesky, scaling, Newton solve for
// the tilt (x*, mu*), psi*.  Device (here), one THREAD per proposal, sequential over k = 1..d:
//     col = L0[k,1:k-1] . Z[1:k-1];  tl = l_k - mu_k - col;  tu = u_k - mu_k - col
//     Z_k = mu_k + TN(0,1; tl, tu)                      (trandn: tail / inverse / rejection regimes)
//     logpr += lnNpr(tl, tu) + mu_k^2/2 - mu_k Z_k
//   accept iff  -log U > psi* - logpr;  output = first nsim accepted proposals in proposal order
//   (ordered compaction as for RSM).  Z is kept coordinate-major (Z[k*P + p]) so that the
//   threads of a warp read and write consecutive addresses; row k of L0 is a broadcast.
// Every uniform is Philox-addressed by (proposal, coordinate, draw) so that the CPU checker can
// replay exactly the same draws.
#include "lqg_common.cuh"
#include "lqg_kernels.h"
#include "lqg_scan.cuh"

namespace lqg {

namespace {

constexpr uint32_t TAG_EXPT = 0x45585054u;
constexpr double kSqrt2 = 1.4142135623730951;

struct Draws {                 // uniform stream of one (proposal, coordinate)
  uint64_t seed, p; uint32_t k, j;
  __device__ __forceinline__ void next2(double& ua, double& ub) {
    uint32_t r[4];
    Philox::gen(j++, (uint32_t)p, (uint32_t)(p >> 32), k, (uint32_t)seed, (uint32_t)(seed >> 32) ^ TAG_EXPT, r);
    ua = u01(r[0], r[1]); ub = u01(r[2], r[3]);
  }
};

// N(0,1) conditional on l < Z < u with l > 0: Rayleigh-proposal rejection (Botev's ntail)
__device__ double ntail(double l, double u, Draws& g) {
  const double c = 0.5 * l * l, f = expm1(c - 0.5 * u * u);
  for (;;) {
    double u1, u2;
    g.next2(u1, u2);
    const double x = c - log(1.0 + u1 * f);
    if (u2 * u2 * x <= c) return sqrt(2.0 * x);
  }
}
// middle regime (Botev's tn): rejection from N(0,1) for wide intervals, inverse transform otherwise
__device__ double tn(double l, double u, Draws& g) {
  if (fabs(u - l) > 2.05) {
    for (;;) {
      double u1, u2;
      g.next2(u1, u2);
      const double x = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
      if (x > l && x < u) return x;
    }
  }
  double u1, u2;
  g.next2(u1, u2);
  const double pl = 0.5 * erfc(l * kInvSqrt2), pu = 0.5 * erfc(u * kInvSqrt2);
  return kSqrt2 * erfcinv(2.0 * (pl - (pl - pu) * u1));
}
__device__ double trandn(double l, double u, Draws& g) {
  const double a = 0.66;
  if (l > a) return ntail(l, u, g);
  if (u < -a) return -ntail(-u, -l, g);
  return tn(l, u, g);
}

}  // namespace

// flags[p] = 1 if accepted; Zbuf[k*P + p]
__global__ void __launch_bounds__(kScanBlock)
expt_propose_kernel(const ExptParams q, int* block_acc) {
  const int64_t p = (int64_t)blockIdx.x * kScanBlock + threadIdx.x;
  const int d = q.d;
  const int64_t P = q.n_proposals;
  int acc = 0;
  if (p < P) {
    const uint64_t gp = (uint64_t)(q.first_proposal + p);
    double logpr = 0.0;
    for (int k = 0; k < d; ++k) {
      const double* lrow = q.Lt + (size_t)k * d;           // row k of L0 (zero diagonal), contiguous
      double col = 0.0;
#pragma unroll 4
      for (int j = 0; j < k; ++j) col = fma(__ldg(lrow + j), q.Z[(size_t)j * P + p], col);
      const double mk = __ldg(q.mu + k);
      const double tl = __ldg(q.l + k) - mk - col, tu = __ldg(q.u + k) - mk - col;
      Draws g{q.seed, gp, (uint32_t)k, 0u};
      const double z = mk + trandn(tl, tu, g);
      q.Z[(size_t)k * P + p] = z;
      logpr += lnNpr(tl, tu) + 0.5 * mk * mk - mk * z;
    }
    Draws g{q.seed, gp, (uint32_t)d, 0u};
    double ua, ub;
    g.next2(ua, ub);
    acc = (-log(ua) > q.psistar - logpr) ? 1 : 0;
    q.flags[p] = (unsigned char)acc;
  }
  const int n = __syncthreads_count(acc);
  if (threadIdx.x == 0) block_acc[blockIdx.x] = n;
}

// ordered compaction: accepted proposal -> output column slot_base + rank (while < nsim)
__global__ void __launch_bounds__(kScanBlock)
expt_emit_kernel(const ExptParams q, const int64_t* acc_off, int64_t slot_base, int64_t nsim,
                 double* out, unsigned long long* counters) {
  __shared__ int wt[kScanBlock / 32];
  const int64_t p = (int64_t)blockIdx.x * kScanBlock + threadIdx.x;
  const int acc = p < q.n_proposals ? q.flags[p] : 0;
  int tot;
  const int local = block_excl_scan(acc, wt, &tot);
  if (!acc) return;
  const int64_t slot = slot_base + acc_off[blockIdx.x] + local;
  if (slot >= nsim) return;
  if (slot == nsim - 1) counters[0] = (unsigned long long)(q.first_proposal + p + 1);
  double* dst = out + (size_t)slot * q.d;
  for (int k = 0; k < q.d; ++k) dst[k] = q.Z[(size_t)k * q.n_proposals + p];
}

// Lt[j + k*d] = L0[k + j*d]
__global__ void expt_transpose_kernel(int d, const double* L0, double* Lt) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)d * d) return;
  const int k = (int)(idx / d), j = (int)(idx % d);
  Lt[idx] = L0[(size_t)k + (size_t)j * d];
}

cudaError_t launch_expt_transpose(int d, const double* L0, double* Lt, cudaStream_t st) {
  const size_t n = (size_t)d * d;
  expt_transpose_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d, L0, Lt);
  ++launch_counter();
  return cudaGetLastError();
}

int expt_blocks(int64_t n) { return (int)((n + kScanBlock - 1) / kScanBlock); }

cudaError_t launch_expt_round(const ExptParams& q, int* block_cnt, int64_t* acc_off, int64_t slot_base,
                              int64_t nsim, double* out, unsigned long long* counters, cudaStream_t st) {
  const int nb = expt_blocks(q.n_proposals);
  expt_propose_kernel<<<nb, kScanBlock, 0, st>>>(q, block_cnt);
  scan_counts_kernel<<<1, 1024, 0, st>>>(block_cnt, acc_off, nb);
  expt_emit_kernel<<<nb, kScanBlock, 0, st>>>(q, acc_off, slot_base, nsim, out, counters);
  launch_counter() += 3;
  return cudaGetLastError();
}

// out[i, col] += v[i]
__global__ void add_colvec_kernel(int m, size_t n, double* out, int64_t ld, const double* v) {
  const size_t total = (size_t)m * n;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx % m);
    out[(idx / m) * ld + i] += v[i];
  }
}
cudaError_t launch_add_colvec(int m, size_t n, double* out, int64_t ld, const double* v, cudaStream_t st) {
  add_colvec_kernel<<<148 * 4, 256, 0, st>>>(m, n, out, ld, v);
  ++launch_counter();
  return cudaGetLastError();
}

}  // namespace lqg
