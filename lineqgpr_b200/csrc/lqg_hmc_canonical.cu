// Exact bouncing HMC in tmg's canonical (whitened) frame with GENERAL dense linear
// constraints F z + g >= 0 — the contract of the reference's only native entry point:
//   .Call("rtmg", n+burn.in, seed, initial2, numlin, f2, g2, 0, NULL)   tmg:R/tmg.R:104
//   rtmg()                       tmg:src/rtmg.cpp:14-66
//   HmcSampler::sampleNext       tmg:src/HmcSampler.cpp:43-133
//   _getNextLinearHitTime        tmg:src/HmcSampler.cpp:152-189
//   _verifyConstraints           tmg:src/HmcSampler.cpp:244-265
//
// Mapping: persistent CTAs; a chain is run by a group of NT threads and NCH groups share one CTA
// (named barrier per group, see Grp in lqg_common.cuh).  The constraint matrix (c x d, column-major
// as R hands it over) is staged ONCE per CTA into shared memory with bulk-async (TMA) copies
// completing on an mbarrier, and then serves every hit search of every chain the CTA's groups
// run: with a small F (cfg1: 199 x 100 = 159 KB) the one staged copy feeds 4 chains at a time
// instead of one.  Hit search = one thread per constraint row (column-major F makes the shared-memory
// reads conflict free, (a_j, b_j) is a 16-byte broadcast), register filter, exact tmg formula
// for the survivors, warp-shuffle arg-min, one slot per warp.  When F does not fit in shared
// memory it is streamed from L2 with the same access pattern.
#include <limits.h>
#include <stdlib.h>

#include <algorithm>

#include "lqg_common.cuh"
#include "lqg_kernels.h"

namespace lqg {

namespace {

constexpr size_t kSmemBudget = 200 * 1024;     // of 227 KB; the rest stays L1
constexpr uint32_t kBulkChunk = 32 * 1024;     // bytes per cp.async.bulk instruction

struct Slot {
  double t; int key; int pad;
};

// deterministic block sum: warp shuffle, one slot per warp, every thread adds the slots in
// the same order.  `buf` has 2*NW entries (double buffered by *parity: one barrier per call).
template <int NW, int NCH>
__device__ __forceinline__ double block_sum(double v, double* buf, int& parity, const Grp gr) {
  v = warp_sum(v);
  if ((gr.tid & 31) == 0) buf[parity * NW + (gr.tid >> 5)] = v;
  group_sync<NCH, NW * 32>(gr.id);
  double s = 0.0;
#pragma unroll
  for (int w = 0; w < NW; ++w) s += buf[parity * NW + w];
  parity ^= 1;
  return s;
}

}  // namespace

template <int NT, int JMAX, int NCH>
__global__ void __launch_bounds__(NT * NCH) hmc_canonical_kernel(const CanonParams p) {
  constexpr int NW = NT / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ Slot slots_[NCH][2][NW];
  __shared__ double sums_[NCH][2 * NW];
  __shared__ int chain_sm_[NCH];
  __shared__ __align__(8) uint64_t mbar;

  Grp gr;
  gr.id = NCH > 1 ? (int)threadIdx.x / NT : 0;
  gr.tid = NCH > 1 ? (int)threadIdx.x - gr.id * NT : (int)threadIdx.x;
  Slot (*slots)[NW] = slots_[gr.id];
  double* sums = sums_[gr.id];
  int& chain_sm = chain_sm_[gr.id];
  auto gsync = [&]() { group_sync<NCH, NT>(gr.id); };
  const int tid = gr.tid, lane = tid & 31, warp = tid >> 5;
  const int d = p.d, c = p.c;

  // smem carve-up: [ab of group 0 .. NCH-1: double2 x d each][F: c*d doubles (optional)]
  const size_t ab_bytes = (((size_t)d * sizeof(double2)) + 127) & ~(size_t)127;
  double2* ab = reinterpret_cast<double2*>(smem_raw + (size_t)gr.id * ab_bytes);
  const double* Fp = p.F;
  if (p.f_in_smem) {
    double* Fs = reinterpret_cast<double*>(smem_raw + (size_t)NCH * ab_bytes);
    const uint32_t total = (uint32_t)((((size_t)c * d * sizeof(double)) + 15) & ~(size_t)15);
    if (threadIdx.x == 0) {
      mbar_init(&mbar, 1);
      mbar_expect_tx(&mbar, total);
      for (uint32_t off = 0; off < total; off += kBulkChunk) {
        const uint32_t n = total - off < kBulkChunk ? total - off : kBulkChunk;
        bulk_g2s(reinterpret_cast<unsigned char*>(Fs) + off,
                 reinterpret_cast<const unsigned char*>(p.F) + off, n, &mbar);
      }
    }
    __syncthreads();
    mbar_wait(&mbar, 0);
    Fp = Fs;
  }

  int sp = 0;   // parity of `sums`
  int hp = 0;   // parity of `slots`

  // work distribution as in the box kernel: group g of CTA b starts with chain g * gridDim.x + b
  bool first_pull = true;
  for (;;) {
    if (tid == 0)
      chain_sm = first_pull ? gr.id * (int)gridDim.x + (int)blockIdx.x : (int)gridDim.x * NCH + atomicAdd(p.next_chain, 1);
    first_pull = false;
    gsync();
    const int chain = chain_sm;
    gsync();
    if (chain >= p.n_chains) break;

    for (int j = tid; j < d; j += NT) ab[j] = make_double2(0.0, p.state[(size_t)chain * d + j]);
    int64_t inj_pos = p.inj_pos ? p.inj_pos[chain] : 0;
    int status = p.chain_status[chain];
    unsigned long long n_bounce = 0, n_search = 0, n_vel = 0, n_chk = 0, n_exact = 0, n_trans = 0;
    gsync();

    for (int s = p.s_begin; s < p.s_end && status == 0; ++s) {
      uint32_t restart = 0;
      for (;;) {                                                   // while(2), :51
        // ---- fresh velocity a ~ N(0, I) (:54-55) ----
        if (p.injected) {
          if ((inj_pos + 1) * d > p.n_injected) { status = LQG_ERR_DRAWS_EXHAUSTED; break; }
          const double* src = p.injected + (size_t)chain * p.n_injected + (size_t)inj_pos * d;
          for (int j = tid; j < d; j += NT) ab[j].x = src[j];
        } else {
          const uint64_t gchain = (uint64_t)(p.chain_offset + chain);
          for (int pr = tid; 2 * pr < d; pr += NT) {
            double n0, n1;
            normal_pair(p.seed, gchain, (uint64_t)inj_pos, (uint32_t)pr, n0, n1);
            ab[2 * pr].x = n0;
            if (2 * pr + 1 < d) ab[2 * pr + 1].x = n1;
          }
        }
        ++inj_pos;                                                 // the chain's draw index
        gsync();
        // energy |a|^2 + |b|^2 is invariant under rotation and reflection: scale of the filter
        double e2 = 0.0;
        for (int j = tid; j < d; j += NT) { const double2 v = ab[j]; e2 = fma(v.x, v.x, fma(v.y, v.y, e2)); }
        const double E = sqrt(block_sum<NW, NCH>(e2, sums, sp, gr));

        double tt = kHalfPi;                                       // :57
        double velsign = 0.0;                                      // :53
        for (;;) {                                                 // while(1), :64
          double S, C;
          sincos(tt, &S, &C);
          // ---- _getNextLinearHitTime (:152-189): one thread per constraint row ----
          double bt = 0.0; int bkey = INT_MAX;
          for (int k = tid; k < c; k += NT) {
            double fa = 0.0, fb = 0.0;
            const double* fk = Fp + k;
#pragma unroll 4
            for (int j = 0; j < d; ++j) {
              const double f = fk[(size_t)j * c];
              const double2 v = ab[j];
              fa = fma(f, v.x, fa);                                // :157
              fb = fma(f, v.y, fb);                                // :158
            }
            const double g = __ldg(p.g + k);
            const double margin = kFilterRel * fma(__ldg(p.fnorm + k), E, fabs(g));
            if (may_hit(fa, fb, g, S, C, margin)) {
              const double t = hit_time_exact(fa, fb, g);
              ++n_exact;
              if (hit_better(t, k, bt, bkey)) { bt = t; bkey = k; }
            }
          }
          Hit w = warp_min_hit(Hit{bt, bkey});
          if (lane == 0) { slots[hp][warp].t = w.t; slots[hp][warp].key = w.key; }
          gsync();
          w.t = 0.0; w.key = INT_MAX;
#pragma unroll
          for (int q = 0; q < NW; ++q) {
            const Slot sl = slots[hp][q];
            if (hit_better(sl.t, sl.key, w.t, w.key)) { w.t = sl.t; w.key = sl.key; }
          }
          hp ^= 1;
          ++n_search;
          const double t = w.t;
          if (t == 0.0 || tt < t) break;                           // :82
          // ---- bounce (:90-110) ----
          if (p.trace_t != nullptr && tid == 0) {
            const int e = p.trace_n[chain];
            if (e < p.trace_cap) { p.trace_t[(size_t)chain * p.trace_cap + e] = t; p.trace_cn[(size_t)chain * p.trace_cap + e] = w.key; }
            p.trace_n[chain] = e + 1;
          }
          ++n_bounce;
          tt = tt - t;                                             // :90
          double st, ct;
          sincos(t, &st, &ct);
          const int cn = w.key;
          double vloc[JMAX], floc[JMAX], nbloc[JMAX];
          double fv = 0.0;
#pragma unroll
          for (int q = 0; q < JMAX; ++q) {
            const int j = tid + q * NT;
            if (j < d) {
              const double2 v = ab[j];
              nbloc[q] = fma(st, v.x, ct * v.y);                   // new_b   (:91)
              vloc[q]  = fma(ct, v.x, -st * v.y);                  // hit_vel (:92)
              floc[q]  = Fp[cn + (size_t)j * c];
              fv = fma(floc[q], vloc[q], fv);
            }
          }
          fv = block_sum<NW, NCH>(fv, sums, sp, gr);
          const double alpha2 = 2.0 * (fv / __ldg(p.frow2 + cn));  // :107-108
          double vs = 0.0;
#pragma unroll
          for (int q = 0; q < JMAX; ++q) {
            const int j = tid + q * NT;
            if (j < d) {
              const double an = fma(-alpha2, floc[q], vloc[q]);    // :109
              vs = fma(an, floc[q], vs);
              ab[j] = make_double2(an, nbloc[q]);
            }
          }
          velsign = block_sum<NW, NCH>(vs, sums, sp, gr);                   // :110 (barrier publishes ab)
          if (velsign < 0.0) break;                                // :112
        }
        if (velsign < 0.0) {                                       // :117
          ++n_vel;
          if (++restart > (uint32_t)p.max_restarts) { status = LQG_ERR_RESTART_CAP; break; }
          continue;
        }
        // ---- last move (:119) into ab[].x, then _verifyConstraints (:244-265) ----
        {
          double S, C;
          sincos(tt, &S, &C);
          for (int j = tid; j < d; j += NT) { const double2 v = ab[j]; ab[j].x = fma(S, v.x, C * v.y); }
        }
        gsync();
        bool fine = true;
        for (int k = tid; k < c; k += NT) {
          double chk = 0.0;
          const double* fk = Fp + k;
#pragma unroll 4
          for (int j = 0; j < d; ++j) chk = fma(fk[(size_t)j * c], ab[j].x, chk);
          fine = fine && (chk + __ldg(p.g + k) >= 0.0);            // :257-258, :121
        }
        if (group_and<NCH, NT>(gr.id, fine)) {
          for (int j = tid; j < d; j += NT) ab[j].y = ab[j].x;     // lastSample = bb (:123)
          gsync();
          break;
        }
        ++n_chk;                                                   // :130
        if (++restart > (uint32_t)p.max_restarts) { status = LQG_ERR_RESTART_CAP; break; }
      }
      if (status != 0) break;
      ++n_trans;
      if (s >= p.burn_in) {
        double* dst = p.out + ((size_t)chain * p.n_per_chain + (s - p.burn_in)) * d;
        for (int j = tid; j < d; j += NT) dst[j] = ab[j].y;
      }
    }

    for (int j = tid; j < d; j += NT) p.state[(size_t)chain * d + j] = ab[j].y;
    n_exact = (unsigned long long)warp_sum((double)n_exact);
    if (lane == 0) atomicAdd(p.stats + 5, n_exact);
    if (tid == 0) {
      if (p.inj_pos) p.inj_pos[chain] = inj_pos;
      p.chain_status[chain] = status;
      atomicAdd(p.stats + 0, n_trans);
      atomicAdd(p.stats + 1, n_bounce);
      atomicAdd(p.stats + 2, n_search);
      atomicAdd(p.stats + 3, n_vel);
      atomicAdd(p.stats + 4, n_chk);
    }
    gsync();
  }
}

// f.f and |f| of every row (tmg recomputes f.f at every bounce, :107)
__global__ void canon_rownorm_kernel(int d, int c, const double* F, double* frow2, double* fnorm) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= c) return;
  double s = 0.0;
  for (int j = 0; j < d; ++j) { const double f = F[k + (size_t)j * c]; s = fma(f, f, s); }
  frow2[k] = s;
  fnorm[k] = sqrt(s);
}

// independent audit: counts (sample, row) pairs with F z + g < -1e-9 (|f||z| + |g|)
__global__ void canon_violation_kernel(int d, int c, size_t n_cols, const double* F,
                                       const double* g, const double* fnorm, const double* out,
                                       unsigned long long* count) {
  unsigned long long bad = 0;
  const size_t total = (size_t)c * n_cols;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(idx % c);
    const double* z = out + (idx / c) * d;
    double s = 0.0, z2 = 0.0;
    for (int j = 0; j < d; ++j) { s = fma(F[k + (size_t)j * c], z[j], s); z2 = fma(z[j], z[j], z2); }
    const double tol = 1e-9 * (fnorm[k] * sqrt(z2) + fabs(g[k]));
    if (!(s + g[k] >= -tol)) ++bad;
  }
  bad = (unsigned long long)warp_sum((double)bad);
  if ((threadIdx.x & 31) == 0 && bad) atomicAdd(count, bad);
}

size_t canon_smem_bytes(int d, int c, int f_in_smem, int nch) {
  size_t ab = (((size_t)d * sizeof(double2)) + 127) & ~(size_t)127;
  size_t f = f_in_smem ? ((((size_t)c * d * sizeof(double)) + 15) & ~(size_t)15) : 0;
  return (size_t)nch * ab + f;
}
int canon_f_fits_smem(int d, int c) { return canon_smem_bytes(d, c, 1, 1) <= kSmemBudget; }

template <int NT, int JMAX, int NCH>
static cudaError_t launch_canon_t(const CanonParams& p, int grid, cudaStream_t st) {
  const size_t smem = canon_smem_bytes(p.d, p.c, p.f_in_smem, NCH);
  auto kern = hmc_canonical_kernel<NT, JMAX, NCH>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (grid <= 0) {
    int per_sm = 0, dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT * NCH, smem);
    if (e != cudaSuccess) return e;
    grid = sms * (per_sm > 0 ? per_sm : 1);
  }
  if ((long long)grid * NCH > p.n_chains) grid = std::max(1, std::min(grid, p.n_chains));
  kern<<<grid, NT * NCH, smem, st>>>(p);
  ++launch_counter();
  return cudaGetLastError();
}

// threads per chain from the problem size, coordinates per thread (JMAX) from d, and — when F is staged
// in shared memory, i.e. one CTA per SM whatever its size — as many chains per CTA as 1024 threads, 128
// registers each and the shared memory left beside F allow (LQG_CANON_NCH=1 forces one chain per CTA)
template <int NT, int NCH>
static cudaError_t launch_canon_j(const CanonParams& p, int grid, cudaStream_t st) {
  const int jm = (p.d + NT - 1) / NT;
  if (jm <= 1) return launch_canon_t<NT, 1, NCH>(p, grid, st);
  if (jm <= 2) return launch_canon_t<NT, 2, NCH>(p, grid, st);
  if (jm <= 4) return launch_canon_t<NT, 4, NCH>(p, grid, st);
  return launch_canon_t<NT, 8, NCH>(p, grid, st);
}
// threads per chain from the problem size, coordinates per thread (JMAX) from d, and — when F is staged
// in shared memory, i.e. one CTA per SM whatever its size — as many chains per CTA as the thread limit
// and the shared memory left beside F allow: 8 chains for up to 128 threads per chain, 2 for 256
// (LQG_CANON_NCH=1 forces one chain per CTA, =4 four)
template <int NT>
static cudaError_t launch_canon_nt(const CanonParams& p, int grid, cudaStream_t st) {
  static const char* env = getenv("LQG_CANON_NCH");
  constexpr int NBIG = NT <= 128 ? 8 : (NT <= 256 ? 2 : 1);
  constexpr int NMID = NT <= 128 ? 4 : NBIG;
  const bool multi = p.f_in_smem && p.n_chains > 1 && !(env && env[0] == '1');
  auto fits = [&](int nch) { return canon_smem_bytes(p.d, p.c, 1, nch) <= kSmemBudget + 16 * 1024; };
  if (multi && NBIG > 1 && fits(NBIG) && !(env && env[0] == '4')) return launch_canon_j<NT, NBIG>(p, grid, st);
  if (multi && NMID > 1 && fits(NMID)) return launch_canon_j<NT, NMID>(p, grid, st);
  return launch_canon_j<NT, 1>(p, grid, st);
}

cudaError_t launch_hmc_canonical(const CanonParams& p, int grid, cudaStream_t st) {
  const int work = p.d > p.c ? p.d : p.c;
  if (p.d <= 64 * 8 && work <= 64)    return launch_canon_nt<64>(p, grid, st);
  if (p.d <= 128 * 8 && work <= 256)  return launch_canon_nt<128>(p, grid, st);
  if (p.d <= 256 * 8)                 return launch_canon_nt<256>(p, grid, st);
  if (p.d <= 1024 * 8)                return launch_canon_nt<1024>(p, grid, st);
  return cudaErrorInvalidValue;
}

cudaError_t launch_canon_rownorms(int d, int c, const double* F, double* frow2, double* fnorm,
                                  cudaStream_t st) {
  canon_rownorm_kernel<<<(c + 127) / 128, 128, 0, st>>>(d, c, F, frow2, fnorm);
  ++launch_counter();
  return cudaGetLastError();
}

cudaError_t launch_canon_violations(int d, int c, size_t n_cols, const double* F, const double* g,
                                    const double* fnorm, const double* out,
                                    unsigned long long* count, cudaStream_t st) {
  canon_violation_kernel<<<148 * 4, 256, 0, st>>>(d, c, n_cols, F, g, fnorm, out, count);
  ++launch_counter();
  return cudaGetLastError();
}

}  // namespace lqg
