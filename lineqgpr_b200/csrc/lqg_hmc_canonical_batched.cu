// Exact bouncing HMC in tmg's canonical frame with a LARGE dense constraint matrix
// (F z + g >= 0, F c x d that does not fit in shared memory) — the rtmg contract
// (tmg:src/rtmg.cpp:14-66, tmg:src/HmcSampler.cpp:43-133,152-189,244-265) at the sizes where the
// reference spends all its time in the dot products f_k.a, f_k.b of every hit search
// (tmg:src/HmcSampler.cpp:157-158: 4 c d flop per bounce).
//
// Lock-step formulation: all chains advance one hit search per round, and the dot products of
// a round are ONE FP64 tensor-core GEMM
//        [F A | F B]  (c x 2n)  =  F (c x d)  x  [A | B] (d x 2n)
// (A, B = velocities and positions of the n chains), so F is read once per round for all
// chains instead of once per chain per bounce (the per-chain kernel of lqg_hmc_canonical.cu is
// L2-bandwidth bound there: 48 MB of F per search at c = 2999, d = 2000).  Per round:
//   start   chains at the start of an attempt draw a ~ N(0, I) (tmg:src/HmcSampler.cpp:54-55),
//           tt = pi/2; b = lastSample at the start of a transition (:48), while a retry keeps b
//           where the failed attempt left it, on its last wall (b is declared outside while(2))
//   GEMM    lqg_dgemm.cu
//   step    one CTA per chain: hit search over the c rows from the GEMM columns (filter + tmg's
//           exact formula, first-index tie break), then either the bounce (:90-110: rotation,
//           reflection with row cn of F, velsign test) or the last move and the feasibility test
//           (:119-130; F bb = sin(tt) F a + cos(tt) F b, the GEMM columns again) and emission.
// Chains that finish early idle until the slowest is done (a few per cent: the number of bounces
// per transition concentrates).
#include <limits.h>

#include "lqg_common.cuh"
#include "lqg_kernels.h"

namespace lqg {

namespace {

constexpr int BN = 256;
constexpr int BW = BN / 32;

struct BSlot { double t; int key; int pad; };

// deterministic block sum (same scheme as the per-chain kernel)
__device__ __forceinline__ double bsum(double v, double* buf, int& parity) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) buf[parity * BW + (threadIdx.x >> 5)] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int w = 0; w < BW; ++w) s += buf[parity * BW + w];
  parity ^= 1;
  return s;
}

__device__ __forceinline__ bool chain_active(const CanonBatch& p, int ch) {
  return p.status[ch] == 0 && p.s_done[ch] < p.total;
}

}  // namespace

__global__ void __launch_bounds__(BN) canon_batch_start_kernel(const CanonBatch p) {
  __shared__ double sums[2 * BW];
  const int ch = blockIdx.x, tid = threadIdx.x;
  if (!chain_active(p, ch) || p.phase[ch] == 1) return;
  const bool fresh = p.phase[ch] == 0;                             // new transition (else: retry)
  const int d = p.d;
  double* a = p.AB + (size_t)ch * d;
  double* b = p.AB + (size_t)(p.n + ch) * d;
  const double* zl = p.Zlast + (size_t)ch * d;
  const int64_t pos = p.pos[ch];
  if (p.injected) {
    if ((pos + 1) * d > p.n_injected) {                            // uniform across the CTA
      if (tid == 0) p.status[ch] = LQG_ERR_DRAWS_EXHAUSTED;
      return;
    }
    const double* src = p.injected + (size_t)ch * p.n_injected + (size_t)pos * d;
    for (int j = tid; j < d; j += BN) a[j] = src[j];
  } else {
    const uint64_t gchain = (uint64_t)(p.chain_offset + ch);
    for (int pr = tid; 2 * pr < d; pr += BN) {
      double n0, n1;
      normal_pair(p.seed, gchain, (uint64_t)pos, (uint32_t)pr, n0, n1);
      a[2 * pr] = n0;
      if (2 * pr + 1 < d) a[2 * pr + 1] = n1;
    }
  }
  __syncthreads();
  double e2 = 0.0;
  for (int j = tid; j < d; j += BN) {
    const double bj = fresh ? zl[j] : b[j], aj = a[j];
    if (fresh) b[j] = bj;                                          // :48; a retry keeps b on its last wall
    e2 = fma(aj, aj, fma(bj, bj, e2));
  }
  int sp = 0;
  const double E = sqrt(bsum(e2, sums, sp));
  if (tid == 0) { p.E[ch] = E; p.tt[ch] = kHalfPi; p.phase[ch] = 1; p.pos[ch] = pos + 1; }
}

__global__ void __launch_bounds__(BN) canon_batch_step_kernel(const CanonBatch p) {
  __shared__ BSlot slots[BW];
  __shared__ double sums[2 * BW];
  const int ch = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (!chain_active(p, ch) || p.phase[ch] != 1) {
    if (tid == 0 && chain_active(p, ch)) atomicAdd(p.n_unfinished, 1);
    return;
  }
  const int d = p.d, c = p.c;
  double* a = p.AB + (size_t)ch * d;
  double* b = p.AB + (size_t)(p.n + ch) * d;
  const double* fa_col = p.FAB + (size_t)ch * c;
  const double* fb_col = p.FAB + (size_t)(p.n + ch) * c;
  const double tt = p.tt[ch], E = p.E[ch];
  int sp = 0;
  double S, C;
  sincos(tt, &S, &C);
  // ---- _getNextLinearHitTime (:152-189) from the GEMM columns ----
  double bt = 0.0; int bkey = INT_MAX;
  unsigned n_exact = 0;
  for (int k = tid; k < c; k += BN) {
    const double fa = fa_col[k], fb = fb_col[k], g = __ldg(p.g + k);
    const double margin = kFilterRel * fma(__ldg(p.fnorm + k), E, fabs(g));
    if (may_hit(fa, fb, g, S, C, margin)) {
      const double t = hit_time_exact(fa, fb, g);
      ++n_exact;
      if (hit_better(t, k, bt, bkey)) { bt = t; bkey = k; }
    }
  }
  Hit w = warp_min_hit(Hit{bt, bkey});
  if (lane == 0) { slots[warp].t = w.t; slots[warp].key = w.key; }
  __syncthreads();
  w.t = 0.0; w.key = INT_MAX;
#pragma unroll
  for (int q = 0; q < BW; ++q) {
    const BSlot sl = slots[q];
    if (hit_better(sl.t, sl.key, w.t, w.key)) { w.t = sl.t; w.key = sl.key; }
  }
  n_exact = (unsigned)warp_sum((double)n_exact);
  if (lane == 0 && n_exact) atomicAdd(p.stats + 5, (unsigned long long)n_exact);
  if (tid == 0) atomicAdd(p.stats + 2, 1ull);                      // hit searches
  const double t = w.t;
  const int s = p.s_done[ch];
  // bookkeeping of a failed attempt (thread 0): returns whether the chain goes on
  auto fail_attempt = [&](int stat_slot) -> bool {
    atomicAdd(p.stats + stat_slot, 1ull);
    const int r = p.restart[ch] + 1;
    p.restart[ch] = r; p.phase[ch] = 2;
    if (r > p.max_restarts) { p.status[ch] = LQG_ERR_RESTART_CAP; return false; }
    return true;
  };
  if (t == 0.0 || tt < t) {                                        // :82 -> last move (:119)
    bool fine = true;
    for (int k = tid; k < c; k += BN) {                            // F bb + g, bb = S a + C b (:121, :244-265)
      const double chk = fma(S, fa_col[k], C * fb_col[k]);
      fine = fine && (chk + __ldg(p.g + k) >= 0.0);
    }
    fine = __syncthreads_and(fine) != 0;
    if (fine) {
      double* zl = p.Zlast + (size_t)ch * d;
      double* dst = (s >= p.burn_in) ? p.out + ((size_t)ch * p.n_per_chain + (s - p.burn_in)) * d : nullptr;
      for (int j = tid; j < d; j += BN) {
        const double bb = fma(S, a[j], C * b[j]);
        zl[j] = bb;                                                // lastSample = bb (:123)
        if (dst) dst[j] = bb;
      }
      if (tid == 0) {
        p.s_done[ch] = s + 1; p.restart[ch] = 0; p.phase[ch] = 0;
        atomicAdd(p.stats + 0, 1ull);
        if (s + 1 < p.total) atomicAdd(p.n_unfinished, 1);
      }
    } else if (tid == 0) {                                         // :130
      if (fail_attempt(4)) atomicAdd(p.n_unfinished, 1);
    }
    return;
  }
  // ---- bounce (:90-110) ----
  double st, ct;
  sincos(t, &st, &ct);
  const int cn = w.key;
  const double* f = p.Ft + (size_t)cn * d;                         // row cn of F, contiguous
  double fv = 0.0;
  for (int j = tid; j < d; j += BN) {
    const double v = fma(ct, a[j], -st * b[j]);                    // hit_vel (:92)
    fv = fma(__ldg(f + j), v, fv);
  }
  fv = bsum(fv, sums, sp);
  const double alpha2 = 2.0 * (fv / __ldg(p.frow2 + cn));          // :107-108
  double vs = 0.0;
  for (int j = tid; j < d; j += BN) {
    const double aj = a[j], bj = b[j], fj = __ldg(f + j);
    const double nb = fma(st, aj, ct * bj);                        // new_b (:91)
    const double v = fma(ct, aj, -st * bj);
    const double an = fma(-alpha2, fj, v);                         // :109
    vs = fma(an, fj, vs);
    a[j] = an; b[j] = nb;
  }
  const double velsign = bsum(vs, sums, sp);                       // :110
  if (tid == 0) {
    atomicAdd(p.stats + 1, 1ull);
    p.tt[ch] = tt - t;                                             // :90
    bool go_on = true;
    if (velsign < 0.0) go_on = fail_attempt(3);                    // :112, :117
    if (go_on) atomicAdd(p.n_unfinished, 1);
  }
}

cudaError_t launch_canon_batch_round_begin(const CanonBatch& p, cudaStream_t st) {
  canon_batch_start_kernel<<<p.n, BN, 0, st>>>(p);
  ++launch_counter();
  return cudaGetLastError();
}
cudaError_t launch_canon_batch_round_step(const CanonBatch& p, cudaStream_t st) {
  canon_batch_step_kernel<<<p.n, BN, 0, st>>>(p);
  ++launch_counter();
  return cudaGetLastError();
}

}  // namespace lqg
