// Exact bouncing HMC for box constraints lb <= eta <= ub, run in the eta frame.
//
// What it replaces: tmvrnorm.HMC -> tmg::rtmg -> HmcSampler::sampleNext
//   R/lineqGPsamplers.R:99-143, tmg:R/tmg.R:1-108, tmg:src/HmcSampler.cpp:43-133,152-189,244-265
//
// tmvrnorm.HMC only ever builds f = [I;-I] (finite rows), g = c(-lb, ub)
// (R/lineqGPsamplers.R:128-136), so every whitened constraint row is +-Ri[i,:] with
// Ri Ri' = Sigma (tmg:R/tmg.R:25-27,50).  Carrying x_a = Ri a and x_b = Ri b instead of (a, b):
//   f_k.a = s x_a[i],  f_k.b = s x_b[i],  f_k.f_k = Sigma[i,i],  g2_k = s(mu_i - bound_i)
//   bounce on coordinate j:  x_b <- sin t x_a + cos t x_b ;  v <- cos t x_a - sin t x_b
//                            x_a <- v - 2 (v_j / Sigma_jj) Sigma[:,j] ;  velsign = s x_a[j]
//   emitted sample = mu + x_b  (no un-whitening GEMM, tmg:R/tmg.R:106)
// so the hit search is O(c) and the reflection O(d) instead of O(c d).
//
// Mapping: one chain per CTA of G warps; every thread keeps CPT coordinates of (x_a, x_b) in
// REGISTERS (coordinate i = tid + k*32G, so the Sigma column and the output are read/written
// coalesced); the per-coordinate filter thresholds sit in registers (small d) or shared memory.
// The bounce-time search is: branch-free register filter -> survivors compacted (ballot/popc)
// into a per-warp shared-memory queue -> tmg's exact formula 32 candidates at a time ->
// integer arg-min with the hardware warp-reduce (REDUX) -> one slot per warp -> one barrier ->
// REDUX over the slots.  CTAs are persistent and pull chains from an atomic queue.
//
// Momenta: a transition attempt (one pass of tmg's while(2) body) consumes ONE fresh momentum
// x_a = U a and ends either in an accepted end point or in a restart; a restart keeps nothing
// but the wall position x_b.  A chain therefore is just a sequential consumer of momenta, like
// tmg's sequential use of nd(eng1) (tmg:src/HmcSampler.cpp:54-55).  In pool mode the momenta of
// the next `attempts` draws of every active chain are produced up front by one FP64
// tensor-core GEMM (lqg_dgemm.cu) and the kernel runs each chain until its pool is used up;
// the host loops rounds until every chain has finished.  Draws are Philox-addressed by
// (chain, draw index), so the samples do not depend on the round size or on the pool.
#include <limits.h>
#include <algorithm>
#include <type_traits>
#include <stdlib.h>
#include <string.h>

#include "lqg_common.cuh"
#include "lqg_kernels.h"

namespace lqg {

// Phase timing of the hit search (diagnostic build only: -DLQG_BOX_TIMING, scripts/probe_box_phases.py):
// lane 0 of warp 0 reads the SM clock at the phase boundaries and the per-phase sums of all chains are
// added into a global table.
#ifdef LQG_BOX_TIMING
__device__ unsigned long long g_box_phase[16];
#define LQG_TICK(idx)                                                   \
  do {                                                                  \
    if (tid == 0) { const long long _t = clock64(); ph_acc[idx] += (unsigned long long)(_t - ph_last); ph_last = _t; } \
  } while (0)
#else
#define LQG_TICK(idx) do { } while (0)
#endif

// Survivors of the filter are compacted into a per-warp queue of wall ids and evaluated 32 at a time.
// A queue entry is just the wall id: the CTA keeps a copy of the chain state (x_a, x_b) and of the
// bound offsets (glo, ghi) in shared memory, so whoever evaluates a wall reads its three operands from
// there — the thread that owns the coordinate never has to move register values through a queue.
constexpr int QCAP = 64;
struct Cand {
  int key[QCAP];
};
// fresh momentum x_a = U a for this chain, computed by the CTA itself (restarts, injected
// draws, or when the GEMM pre-pass is off).  a ~ N(0, I_d) drawn like tmg:src/HmcSampler.cpp:54-55.
// `xs` (shared, CPT*T doubles) first stages a, then receives x_a.  Out of line on purpose: its
// accumulators must not count against the register budget of the hit-search loop.
template <int G, int CPT, int NCH>
static __device__ __noinline__ bool draw_momentum(const BoxParams& p, int chain, int64_t draw, double* xs, const Grp gr) {
  constexpr int T = 32 * G;
  const int tid = gr.tid;
  const int d = p.d;
  double* a_sm = xs;
  group_sync<NCH, T>(gr.id);                                       // nobody reads the old x_a copy any more
  if (p.injected) {
    if ((draw + 1) * d > p.n_injected) return false;              // uniform across the CTA
    const double* src = p.injected + (size_t)chain * p.n_injected + (size_t)draw * d;
    for (int i = tid; i < d; i += T) a_sm[i] = src[i];
  } else {
    const uint64_t gchain = (uint64_t)(p.chain_offset + (int64_t)chain * p.chain_stride);
    for (int pr = tid; 2 * pr < d; pr += T) {
      double n0, n1;
      normal_pair(p.seed, gchain, (uint64_t)draw, (uint32_t)pr, n0, n1);
      a_sm[2 * pr] = n0;
      if (2 * pr + 1 < d) a_sm[2 * pr + 1] = n1;
    }
  }
  group_sync<NCH, T>(gr.id);
  double acc[CPT];
#pragma unroll
  for (int k = 0; k < CPT; ++k) acc[k] = 0.0;
  const double* U = p.U;
  // x_a[i] = sum_col U[i,col] a[col]; U column-major so a warp reads 256 contiguous bytes.
  // Loads are unconditional (row index clamped) so that CB columns x CPT rows are in flight
  // per thread; for an upper-triangular U the slabs k with k*T > col are all zero and are
  // skipped with a CTA-uniform test.
  constexpr int CB = (CPT >= 16) ? 2 : 4;
  int rows[CPT];
#pragma unroll
  for (int k = 0; k < CPT; ++k) rows[k] = min(tid + k * T, d - 1);
  for (int col0 = 0; col0 < d; col0 += CB) {
    double u[CB][CPT];
#pragma unroll
    for (int c = 0; c < CB; ++c) {
      const int col = min(col0 + c, d - 1);
      const double* ucol = U + (size_t)col * d;
#pragma unroll
      for (int k = 0; k < CPT; ++k)
        u[c][k] = (!p.u_upper || k * T <= col0 + CB - 1) ? __ldg(ucol + rows[k]) : 0.0;
    }
#pragma unroll
    for (int c = 0; c < CB; ++c) {
      const double av = (col0 + c < d) ? a_sm[col0 + c] : 0.0;
#pragma unroll
      for (int k = 0; k < CPT; ++k) acc[k] = fma(u[c][k], av, acc[k]);
    }
  }
  group_sync<NCH, T>(gr.id);                                                 // a is consumed: the area now receives x_a
#pragma unroll
  for (int k = 0; k < CPT; ++k) xs[tid + k * T] = (tid + k * T < d) ? acc[k] : 0.0;
  group_sync<NCH, T>(gr.id);
  return true;
}

// Search windows.  A sampleNext step needs the FIRST wall hit; with ~16 bounces per transition at m = 1000
// the walls that could be hit anywhere in the remaining time are ~100, the ones hit within a few mean gaps
// a handful.  The fast search therefore looks at a window [0, tau] first, tau = (pi/2) 2^-l chosen from the
// recent hit times of the transition (4 x their running mean): the filter runs with (sin tau, cos tau) and a
// hit is accepted only strictly inside the window (tan(t/2) <= tan(tau/2) (1 - 4e-6), so that a wall the
// window filter may have dropped is farther from the winner than the near-tie band that sends a search to
// tmg's arithmetic).  A window without a hit is widened 4 x and searched again; a window that reaches the
// remaining time IS the remaining-time search.  The policy depends on the chain's own hit times only, so
// samples stay independent of kernel shape, round size and sharding.
constexpr int kNLV = 12;
__constant__ double c_winK[kNLV] = {1.0, 0.41421356237309503, 0.198912367379658, 0.09849140335716425, 0.049126849769467254,
                                    0.024548622108925444, 0.012272462379566276, 0.006136000157623402, 0.003067971201422665,
                                    0.0015339819910886664, 0.0007669905443430926, 0.000383495215771441};     // tan(tau/2)
__constant__ double c_winS[kNLV] = {1.0, 0.7071067811865476, 0.3826834323650898, 0.19509032201612828, 0.0980171403295606,
                                    0.049067674327418015, 0.024541228522912288, 0.012271538285719925, 0.006135884649154475,
                                    0.003067956762965976, 0.0015339801862847657, 0.0007669903187427045};     // sin tau
__constant__ double c_winC[kNLV] = {6.123233995736766e-17, 0.7071067811865476, 0.9238795325112867, 0.9807852804032304,
                                    0.9951847266721969, 0.9987954562051724, 0.9996988186962042, 0.9999247018391445,
                                    0.9999811752826011, 0.9999952938095762, 0.9999988234517019, 0.9999997058628822};  // cos tau
constexpr double kWinKavg0 = 0.05;     // running mean of tan(t/2) at the start of a transition (~pi/32 between hits)

// One slot per warp for the cross-warp stage of the arg-min.  `tb` orders the candidates: bits of the
// hit time (exact search) or of tan(t/2) (fast search); hi2 = high word of the warp's runner-up.
struct WinSlot {
  unsigned long long tb; double xa, xb, st, ct; int key; unsigned hi2; int slow; int pad;
};
// Bounces whose exact hit time is still owed to tt (see hit_fast): wall id and the (f.a, f.b) it was hit with
constexpr int DCAP = 128;
struct Deferred {
  double fa[DCAP]; double fb[DCAP]; int key[DCAP];
};

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// What the rarely taken paths need from the kernel: all of it lives in shared or global memory, so
// they can be out-of-line functions that do not compete for the registers of the search loop.
struct BoxShared {
  const double* xa_s; const double* xb_s;         // state copy, padded to NS entries
  const double* glo; const double* ghi;            // bound offsets: shared copy when available, else global
  int d, ns;
};

// One hit search with tmg's own arithmetic (tmg:src/HmcSampler.cpp:152-189): every wall is evaluated
// with hit_time_exact, arg-min with the lowest index winning ties.  Taken when the fast search meets
// a wall on one of tmg's thresholds, a near-tie, or a hit within 1e-9 of the remaining time, and for
// every search with lqg_opts.hmc_search = 1.  The winner is left in red[0].
template <int G, int NCH>
static __device__ __noinline__ void exact_search(const BoxShared sh, WinSlot* red, unsigned* n_eval, const Grp gr) {
  constexpr int T = 32 * G;
  const int tid = gr.tid, lane = tid & 31, warp = tid >> 5;
  const int d = sh.d;
  unsigned long long btb = kNoHitBits; int bkey = INT_MAX;
  double bxa = 0.0, bxb = 0.0;
  unsigned cnt = 0;
  for (int key = tid; key < 2 * d; key += T) {
    const int i = key < d ? key : key - d;
    const double g = key < d ? sh.glo[i] : sh.ghi[i];
    if (g >= 0.5 * kBigG) continue;                               // dropped (infinite) bound: no row in f
    const double sgn = key < d ? 1.0 : -1.0;
    const double a_ = sh.xa_s[i], b_ = sh.xb_s[i];
    const unsigned long long tb = hit_bits(hit_time_exact(sgn * a_, sgn * b_, g));
    ++cnt;
    if (tb < btb || (tb == btb && key < bkey)) { btb = tb; bkey = key; bxa = a_; bxb = b_; }
  }
  unsigned long long mtb; int mkey;
  const int src = warp_argmin_hit(btb, bkey, mtb, mkey);
  group_sync<NCH, T>(gr.id);                                                // red[] may still be read from the fast search
  if (lane == src) { WinSlot& r = red[warp]; r.tb = mtb; r.key = mkey; r.xa = bxa; r.xb = bxb; }
  group_sync<NCH, T>(gr.id);
  if (tid == 0) {
    int ws = 0;
    for (int w = 1; w < G; ++w)
      if (red[w].tb < red[ws].tb || (red[w].tb == red[ws].tb && red[w].key < red[ws].key)) ws = w;
    if (ws != 0) red[0] = red[ws];
  }
  *n_eval += cnt;
  group_sync<NCH, T>(gr.id);
}

// exact hit times of the bounces still owed to tt (see hit_fast), one lane per bounce; the caller
// subtracts them in order
template <int G, int NCH>
static __device__ __noinline__ void settle_times(const BoxShared sh, Deferred* dl, int n_dl, const Grp gr) {
  constexpr int T = 32 * G;
  group_sync<NCH, T>(gr.id);                                      // entries written by thread 0 are visible
  for (int e = gr.tid; e < n_dl; e += T) {
    const int key = dl->key[e];
    const double g = key < sh.d ? sh.glo[key] : sh.ghi[key - sh.d];
    dl->fa[e] = hit_time_exact(dl->fa[e], dl->fb[e], g);
  }
  group_sync<NCH, T>(gr.id);
}

// THR: where the filter thresholds live: 0 = registers, 1 = shared memory, 2 = recomputed from the
//      bound offsets.   GS: bound offsets (glo, ghi) in shared memory (else read through L1).
//  RS: the thread also keeps its coordinates of (x_a, x_b) in registers (the filter and the update then
//      work from registers and only mirror the result to shared memory); RS = false trades those 4 CPT
//      registers for shared-memory traffic so that one more CTA fits on the SM.
//  NCH: chains per CTA (groups of G warps, see Grp).  Fewer warps per chain and more chains per SM: the
//      per-search work that every warp of a chain repeats (slot scan, decisions, time bookkeeping) and the
//      barriers shrink with G, and the chains of an SM hide each other's latency; the per-coordinate tables
//      are shared by the groups, so a chain costs 2 NS doubles of shared memory.
template <int G, int CPT, int THR, bool GS, int MINB, bool PACK = false, bool RS = true, int NCH = 1>
__global__ void __launch_bounds__(32 * G * NCH, MINB) hmc_box_kernel(const BoxParams p) {
  constexpr int T = 32 * G;
  constexpr int NS = CPT * T;                      // padded dimension: one slot per (thread, k)
  static_assert(NCH >= 1 && NCH <= 15, "one named barrier per chain group");
  static_assert(CPT <= 32, "two filter bits per coordinate in a 64-bit mask");
  extern __shared__ double dyn_sm[];
  Grp gr;
  gr.id = NCH > 1 ? (int)threadIdx.x / T : 0;
  gr.tid = NCH > 1 ? (int)threadIdx.x - gr.id * T : (int)threadIdx.x;
  // tables shared by the groups, then per group the state copy
  double* glo_s = dyn_sm;                          // [NS] if GS
  double* ghi_s = glo_s + NS;
  double* hlo_s = GS ? ghi_s + NS : glo_s;         // [NS] if THR == 1
  double* hhi_s = hlo_s + NS;
  double* st_base = dyn_sm + (size_t)NS * ((GS ? 2 : 0) + (THR == 1 ? 2 : 0));
  // xa_s doubles as the staging area of a fresh draw (draw_momentum): x_a is being replaced then
  double* xa_s = st_base + (size_t)gr.id * 2 * NS; // [NS] copy of x_a   (the two copies swap roles when a transition
  double* xb_s = xa_s + NS;                        // [NS] copy of x_b    is accepted, see the end of the attempt loop)
  __shared__ WinSlot red_[NCH][2][G];
  __shared__ Cand queue_[NCH][G];
  __shared__ int qcnt_[NCH][G];                    // PACK: candidates left in every warp's queue after the filter
  __shared__ Deferred dl_[NCH];
  __shared__ int chain_sm_[NCH];
  WinSlot (*red)[G] = red_[gr.id];
  Cand* queue = queue_[gr.id];
  int* qcnt = qcnt_[gr.id];
  Deferred& dl = dl_[gr.id];
  int& chain_sm = chain_sm_[gr.id];
  auto gsync = [&]() { group_sync<NCH, T>(gr.id); };

  const int tid = gr.tid, lane = tid & 31, warp = tid >> 5;
  const int d = p.d;
  const bool fast_on = p.fast != 0;

  // Per-coordinate filter thresholds h = g - m, with g the bound offset (mu - lb or ub - mu) and
  // m = 3e-9 |g| + abs_eps the safety margin (orders of magnitude above any rounding difference
  // between this eta-frame recursion and tmg's whitened dot products).
  constexpr int NG = THR == 0 ? CPT : 1;
  double hlo_r[NG], hhi_r[NG];
  const double abs_eps = p.abs_eps;
  auto thr = [&](double g) { return g - fma(3.0 * kFilterRel, fabs(g), abs_eps); };
  for (int i = threadIdx.x; i < NS; i += T * NCH) {
    const double gl = i < d ? p.glo[i] : kBigG, gh = i < d ? p.ghi[i] : kBigG;
    if (GS) { glo_s[i] = gl; ghi_s[i] = gh; }
    if (THR == 1) { hlo_s[i] = thr(gl); hhi_s[i] = thr(gh); }
  }
  if (THR == 0) {
#pragma unroll
    for (int k = 0; k < NG; ++k) {
      const int i = tid + k * T;
      hlo_r[k] = thr(i < d ? p.glo[i] : kBigG);
      hhi_r[k] = thr(i < d ? p.ghi[i] : kBigG);
    }
  }
  __syncthreads();
  BoxShared sh;
  sh.xa_s = xa_s; sh.xb_s = xb_s; sh.glo = GS ? glo_s : p.glo; sh.ghi = GS ? ghi_s : p.ghi; sh.d = d; sh.ns = NS;
  auto g_of = [&](int key) -> double {             // bound offset of wall `key` (lower walls 0..d-1, upper d..2d-1)
    if (GS) return key < d ? glo_s[key] : ghi_s[key - d];
    return __ldg(key < d ? p.glo + key : p.ghi + (key - d));
  };

  const int total = p.burn_in + p.n_per_chain;
  // Work distribution: group g of CTA b starts with entry g * gridDim.x + b of the work list (a list shorter
  // than the chain slots is spread over the SMs first), later entries come from the atomic queue.
  bool first_pull = true;
  for (;;) {
    if (tid == 0)
      chain_sm = first_pull ? gr.id * (int)gridDim.x + (int)blockIdx.x : (int)gridDim.x * NCH + atomicAdd(p.next_chain, 1);
    first_pull = false;
    gsync();
    const int aidx = chain_sm;
    gsync();
    if (aidx >= p.n_active) break;
    const int chain = p.active ? p.active[aidx] : aidx;

    double xa[RS ? CPT : 1], xb[RS ? CPT : 1];
    auto XA = [&](int k) -> double { return RS ? xa[RS ? k : 0] : xa_s[tid + k * T]; };
    auto XB = [&](int k) -> double { return RS ? xb[RS ? k : 0] : xb_s[tid + k * T]; };
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
      const int i = tid + k * T;
      const double b0 = i < d ? p.state[(size_t)chain * d + i] : 0.0;
      if (RS) { xb[RS ? k : 0] = b0; xa[RS ? k : 0] = 0.0; }
      xb_s[i] = b0;
    }
    int s = p.s_done[chain];                       // transitions completed so far
    int64_t draw = p.pos[chain];                   // momenta consumed so far
    int rcount = p.rcount[chain];                  // restarts inside the current transition
    int status = p.chain_status[chain];
    unsigned n_bounce = 0, n_search = 0, n_vel = 0, n_chk = 0, n_exact = 0, n_trans = 0, n_slow = 0, n_widen = 0;
    int parity = 0;
#ifdef LQG_BOX_TIMING
    unsigned long long ph_acc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long ph_last = clock64();
#endif

    // this chain's momenta of the round: pool columns [col0, col0 + my_att)
    const size_t col0 = p.pool == nullptr ? 0 : p.col_off ? (size_t)p.col_off[aidx] : (size_t)aidx * p.attempts;
    const int my_att = (p.pool != nullptr && p.col_off) ? p.col_off[aidx + 1] - (int)col0 : p.attempts;
    for (int att = 0; att < my_att && s < total && status == 0; ++att) {   // while(2), HmcSampler.cpp:51
      if (p.pool != nullptr) {
        const double* src = p.pool + (col0 + att) * d;
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
          const int i = tid + k * T;
          const double a0 = i < d ? __ldcs(src + i) : 0.0;       // streamed once: keep Sigma in L2 instead
          if (RS) xa[RS ? k : 0] = a0;
          xa_s[i] = a0;
        }
        // the next momentum of this chain is on its way to L2 while this attempt runs (one request per line)
        if (att + 1 < my_att && (tid & 15) == 0) {
#pragma unroll
          for (int k = 0; k < CPT; ++k) {
            const int i = tid + k * T;
            if (i < d) prefetch_l2(src + d + i);
          }
        }
      } else {
        if (!draw_momentum<G, CPT, NCH>(p, chain, draw, xa_s, gr)) {
          status = LQG_ERR_DRAWS_EXHAUSTED;
          break;
        }
#pragma unroll
        for (int k = 0; k < CPT; ++k) if (RS) xa[RS ? k : 0] = xa_s[tid + k * T];
      }
      ++draw;
      LQG_TICK(11);                                              // 11: a fresh momentum is in place (pool column or in-kernel draw)
      double tt = kHalfPi;                                       // :57; exact up to the hit times still owed (n_dl)
      double velsign = 0.0;                                      // :53
      // sin/cos of the remaining time, advanced by angle subtraction after every bounce.  They feed
      // the filter, the fast search's comparison with the remaining time (both with margins) and,
      // after fast bounces, the final rotation; settling re-bases them on the exact tt.
      double S = 1.0, C = 6.123233995736766e-17;                 // sin(M_PI/2), cos(M_PI/2)
      int n_dl = 0;                                              // bounces whose exact t is still owed to tt
      double kavg = kWinKavg0;                                   // running mean of tan(t/2) of this transition's hits
      int wlvl = 3;                                              // current window level (see c_winK)

      // tt -= t for every owed bounce, in order, with tmg's own hit-time formula
      auto settle = [&]() {
        if (n_dl == 0) return;
        settle_times<G, NCH>(sh, &dl, n_dl, gr);
        for (int e = 0; e < n_dl; ++e) tt = tt - dl.fa[e];       // :90, same order, same operands
        n_dl = 0;
        sincos(tt, &S, &C);
        gsync();                                         // dl is reused by the next bounce
      };

      // The fast search's filter result (two bits per coordinate of this thread) and the window it was
      // computed for.  After a bounce the new (x_a, x_b) are in registers inside the update loop and the
      // sine / cosine of the new remaining time are known, so the filter of the NEXT search is evaluated
      // right there (one pass over the state per bounce instead of two); a separate filter pass only runs
      // for the first search of an attempt and when a window held no hit.
      using Mask = typename std::conditional<(CPT > 16), unsigned long long, unsigned>::type;
      Mask mask = 0;
      bool have_mask = false;
      double ktt = 0.0, klim = 0.0, Sf = S, Cf = C;
      bool win_full = true;
      // window of the coming search (see c_winK): the smallest dyadic angle whose tan(tau/2) covers 4 x the
      // running mean of the hit times; from the current (S, C) = sin / cos of the remaining time
      auto select_window = [&]() {
        ktt = S / (1.0 + C);                                     // tan(tt/2) of the remaining time
        win_full = !p.window;
        Sf = S; Cf = C;
        klim = fma(ktt, 1e-9, ktt) + 1e-12;
        if (!win_full) {
          const double target = 4.0 * kavg;
          while (wlvl + 1 < kNLV && c_winK[wlvl + 1] >= target) ++wlvl;
          while (wlvl > 0 && c_winK[wlvl] < target) --wlvl;
          const double kwin = c_winK[wlvl];
          win_full = kwin >= 0.999 * ktt;                        // reaches the remaining time: search all of it
          if (!win_full) { Sf = c_winS[wlvl]; Cf = c_winC[wlvl]; klim = kwin * (1.0 - 4e-6); }
        }
      };
      // Filter of coordinate slot k (see may_hit in lqg_common.cuh).  With y = position at the end of the
      // window, dy = velocity there, u^2 = a^2 + b^2 and the threshold h = g - margin:
      //   lower row (f = +e_i):  safe iff min(b, y) > -h, unless the orbit dips in between
      //                          (a < 0 and dy > 0) AND reaches the wall (u^2 >= h^2)
      //   upper row (f = -e_i):  mirrored.
      // h |h| compares like h^2 for h > 0 and is negative (always reachable) for h <= 0.  Straight-line
      // predicates (bitwise & |: no short-circuit branches, no fmin/fmax NaN fix-ups).
      auto filter_bits = [&](int k, double a_, double b_) -> Mask {
        const int i = tid + k * T;
        const double y = fma(a_, Sf, b_ * Cf);
        const double dy = fma(a_, Cf, -b_ * Sf);
        const double u2 = fma(a_, a_, b_ * b_);
        double hl, hh;
        if (THR == 0) { hl = hlo_r[THR == 0 ? k : 0]; hh = hhi_r[THR == 0 ? k : 0]; }
        else if (THR == 1) { hl = hlo_s[i]; hh = hhi_s[i]; }
        else { hl = thr(g_of(i < d ? i : 0)); hh = thr(g_of(i < d ? d + i : d)); if (i >= d) { hl = kBigG; hh = kBigG; } }
        const double nhl = -hl;
        const bool safe_lo = (b_ > nhl) & (y > nhl);
        const bool safe_hi = (b_ < hh) & (y < hh);
        const bool dip_lo = (a_ < 0.0) & (dy > 0.0) & (u2 >= hl * fabs(hl));
        const bool dip_hi = (a_ > 0.0) & (dy < 0.0) & (u2 >= hh * fabs(hh));
        const bool need_lo = !safe_lo | dip_lo;
        const bool need_hi = !safe_hi | dip_hi;
        return ((Mask)(need_lo ? 1u : 0u) << (2 * k)) | ((Mask)(need_hi ? 1u : 0u) << (2 * k + 1));
      };

      for (;;) {                                                 // while(1), :64
        unsigned long long wtb = kNoHitBits; int wkey = INT_MAX;
        double wxa = 0.0, wxb = 0.0, wst = 0.0, wct = 0.0;
        bool need_exact = !fast_on;
        if (fast_on) for (;;) {
          // ---- fast hit search: walls that survive the filter are ranked by tan(t/2) (hit_fast) ----
          // lane state: best candidate, high word of its runner-up, "some wall needs tmg's arithmetic"
          unsigned long long btb = kNoHitBits; int bkey = INT_MAX;
          double bxa = 0.0, bxb = 0.0, bst = 0.0, bct = 0.0;
          unsigned bhi2 = 0xffffffffu;
          bool bslow = false;
          if (!have_mask) select_window();                       // else: chosen when the mask was computed (bounce update)
          auto consider = [&](int key) {
            const int i = key < d ? key : key - d;
            const double sgn = key < d ? 1.0 : -1.0;             // wall row f = sgn e_i
            const double a_ = xa_s[i], b_ = xb_s[i];
            const double g = g_of(key);
            FastHit h;
            h.st = 0.0; h.ct = 0.0;
            const int cls = hit_fast(sgn * a_, sgn * b_, g, klim, h);
            bslow |= cls == kFastSlow;
            const unsigned long long tb = cls == kFastHit ? (unsigned long long)__double_as_longlong(h.key) : kNoHitBits;
            if (tb < btb || (tb == btb && key < bkey)) {
              bhi2 = min(bhi2, (unsigned)(btb >> 32));
              btb = tb; bkey = key; bxa = a_; bxb = b_; bst = h.st; bct = h.ct;
            } else {
              bhi2 = min(bhi2, (unsigned)(tb >> 32));
            }
          };
          Cand* q = &queue[warp];
          int qn = 0;                                            // warp-uniform queue length
          auto flush = [&]() {
            __syncwarp();                                        // queue writes visible to every lane
            for (int e = lane; e < qn; e += 32) consider(q->key[e]);
            __syncwarp();                                        // entries consumed before they are overwritten
            n_exact += (lane == 0) ? qn : 0;
            qn = 0;
          };
          LQG_TICK(0);                                           // 0: everything since the last bounce update
          if (!have_mask) {
            // separate filter pass: straight-line code over the thread's coordinates (no votes, no branches: the
            // coordinates overlap in the pipeline); with the state in shared memory nothing forces a full
            // unroll: 8 coordinates in flight are enough
            mask = 0;
#pragma unroll (RS ? CPT : (CPT < 8 ? CPT : 8))
            for (int k = 0; k < CPT; ++k) mask |= filter_bits(k, XA(k), XB(k));
          }
          have_mask = false;
          LQG_TICK(1);                                           // 1: filter masks
          // Phase 2: the survivors' wall ids are compacted into the warp's queue: exclusive prefix sum of
          // the per-lane counts, then every lane walks its set bits (most lanes have none).
          const int cnt = (CPT > 16) ? __popcll((unsigned long long)mask) : __popc((unsigned)mask);
          int incl = cnt;
#pragma unroll
          for (int off = 1; off < 32; off <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, off);
            incl += lane >= off ? v : 0;
          }
          const int wtotal = __shfl_sync(0xffffffffu, incl, 31);
          // more survivors than queue slots (tight boxes in the one- and two-warp variants): rounds of
          // QCAP walls, each evaluated by the warp itself before the next round overwrites the queue
          for (int base = 0; base < wtotal; base += QCAP) {
            int pos = incl - cnt - base;
            for (Mask mm = mask; mm != 0; mm &= mm - 1) {
              const int b = (CPT > 16) ? __ffsll((long long)mm) - 1 : __ffs((int)mm) - 1;
              if ((unsigned)pos < (unsigned)QCAP) q->key[pos] = ((b & 1) ? d : 0) + tid + (b >> 1) * T;
              ++pos;
            }
            qn = min(QCAP, wtotal - base);
            if (wtotal > QCAP) flush();
          }
          LQG_TICK(2);                                           // 2: compaction
          if (PACK && G > 1) {
            // Candidates of all warps are evaluated together, 32 per warp from the front: with ~13 per
            // warp, 8 partly filled passes become ~4 full ones.
            if (lane == 0) qcnt[warp] = qn;
            gsync();
            LQG_TICK(3);                                         // 3: barrier after the pushes
            int ntot = 0;
#pragma unroll
            for (int w = 0; w < G; ++w) ntot += qcnt[w];
            for (int e = tid; e < ntot; e += T) {
              int w = 0, le = e;
#pragma unroll
              for (int u = 0; u < G - 1; ++u) {                  // queue that holds entry e (static indexing only)
                const int cw = qcnt[u];
                const bool past = (w == u) & (le >= cw);
                le -= past ? cw : 0;
                w += past ? 1 : 0;
              }
              consider(queue[w].key[le]);
            }
            if (tid == 0) n_exact += ntot;
            // the slot exchange below has its own barrier, which also protects the queues from the
            // next search's pushes
          } else {
            flush();
          }
          LQG_TICK(4);                                           // 4: candidate evaluation
          // ---- arg-min: hardware warp-reduce, the owner lane fills the warp's slot, then every thread
          //      scans the G slots ----
          unsigned long long mtb; int mkey;
          const int src = warp_argmin_hit(btb, bkey, mtb, mkey);
          unsigned whi2 = __reduce_min_sync(0xffffffffu, lane == src ? bhi2 : (unsigned)(btb >> 32));
          bool wslow = __any_sync(0xffffffffu, bslow);
          if (G > 1) {
            if (lane == src) {
              WinSlot& r = red[parity][warp];
              r.tb = mtb; r.key = mkey; r.xa = bxa; r.xb = bxb; r.hi2 = whi2; r.slow = wslow ? 1 : 0;
              r.st = bst; r.ct = bct;
            }
            LQG_TICK(5);                                         // 5: warp arg-min + slot
            gsync();
            LQG_TICK(6);                                         // 6: barrier of the slot exchange
            int ws = 0;
            wtb = red[parity][0].tb; wkey = red[parity][0].key;
            whi2 = 0xffffffffu;
            int anyslow = red[parity][0].slow;
#pragma unroll
            for (int w = 1; w < G; ++w) {                        // broadcast reads: G is at most 16
              const unsigned long long tb_w = red[parity][w].tb;
              const int key_w = red[parity][w].key;
              anyslow |= red[parity][w].slow;
              if (tb_w < wtb || (tb_w == wtb && key_w < wkey)) {
                whi2 = min(whi2, (unsigned)(wtb >> 32));
                wtb = tb_w; wkey = key_w; ws = w;
              } else {
                whi2 = min(whi2, (unsigned)(tb_w >> 32));
              }
            }
            const WinSlot& r = red[parity][ws];
            wxa = r.xa; wxb = r.xb; wst = r.st; wct = r.ct;
            whi2 = min(whi2, r.hi2); wslow = anyslow != 0;
            parity ^= 1;
          } else {
            wtb = mtb; wkey = mkey;
            wxa = __shfl_sync(0xffffffffu, bxa, src);
            wxb = __shfl_sync(0xffffffffu, bxb, src);
            wst = __shfl_sync(0xffffffffu, bst, src);
            wct = __shfl_sync(0xffffffffu, bct, src);
          }
          // tmg's arithmetic decides whenever a wall sits on one of its thresholds, two walls are hit
          // within ~1e-6 relative of each other, or the first hit is within 1e-9 of the remaining time
          const bool tie = wtb != kNoHitBits && whi2 <= (unsigned)(wtb >> 32) + 1u;
          bool close = false;
          if (wtb != kNoHitBits) {
            const double kw = __longlong_as_double((long long)wtb);
            close = fabs(kw - ktt) <= fma(ktt, 1e-9, 1e-12);
          }
          need_exact = wslow | tie | close;
          if (win_full || need_exact || wtb != kNoHitBits) {
            if (!need_exact && wtb != kNoHitBits) kavg = 0.5 * (kavg + __longlong_as_double((long long)wtb));
            break;
          }
          kavg *= 4.0;                                           // nothing inside the window: widen and search again
          ++n_widen;
        }
        LQG_TICK(7);                                             // 7: slot scan + decision
        double st, ct;
        if (need_exact) {
          if (fast_on) ++n_slow;
          settle();                                              // the exact comparison tt < t needs the exact tt
          exact_search<G, NCH>(sh, red[parity], &n_exact, gr);
          const WinSlot& r = red[parity][0];
          wtb = r.tb; wkey = r.key; wxa = r.xa; wxb = r.xb;
          gsync();                                       // the slot is free for the next search
          ++n_search;
          const double t = hit_from_bits(wtb);
          if (t == 0.0 || tt < t) break;                         // :82
          tt = tt - t;                                           // :90
          sincos(t, &st, &ct);
        } else {
          ++n_search;
          if (wtb == kNoHitBits) break;                          // no wall within the remaining time (:82)
          st = wst; ct = wct;
          if (n_dl == DCAP) settle();
          if (tid == 0) {                                        // the exact t of this hit is owed to tt
            const double sgn = wkey < d ? 1.0 : -1.0;
            dl.fa[n_dl] = sgn * wxa; dl.fb[n_dl] = sgn * wxb; dl.key[n_dl] = wkey;
          }
          ++n_dl;
        }
        // ---- bounce: rotate to the wall, reflect (:90-110) ----
        ++n_bounce;
        const int j = wkey < d ? wkey : wkey - d;
        const double side = wkey < d ? 1.0 : -1.0;
        const double* scol = p.Sigma + (size_t)j * d;
        double sc[CPT];
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
          const int i = tid + k * T;
          sc[k] = i < d ? __ldg(scol + i) : 0.0;
        }
        const double sjj = __ldg(p.sdiag + j);
#ifdef LQG_BOX_TIMING
        if (tid == 0) { double acc_ = 0; _Pragma("unroll") for (int k = 0; k < CPT; ++k) acc_ += sc[k]; if (acc_ == 1.2345e300) ph_acc[15] += 1; }
        LQG_TICK(8);                                             // 8: Sigma column arrives (L2 latency)
#endif
        const double vj = __dsub_rn(__dmul_rn(ct, wxa), __dmul_rn(st, wxb));   // hit_vel[j]
        const double alpha2 = 2.0 * (vj / sjj);                  // 2*alpha, alpha = f.v / f.f
        const double Sn = fma(S, ct, -C * st), Cn = fma(C, ct, S * st);   // sin/cos(tt - t)
        S = Sn; C = Cn;
        if (fast_on) select_window();                            // window of the next search: its filter runs in the loop below
        mask = 0;
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
          const double a_ = XA(k), b_ = XB(k);
          const double nb = fma(st, a_, ct * b_);                // new_b  (:91)
          const double v = fma(ct, a_, -st * b_);                // hit_vel (:92)
          const double na = fma(-alpha2, sc[k], v);              // reflected velocity (:109)
          if (RS) { xa[RS ? k : 0] = na; xb[RS ? k : 0] = nb; }
          xa_s[tid + k * T] = na; xb_s[tid + k * T] = nb;        // the copy the next search evaluates from
          if (fast_on) mask |= filter_bits(k, na, nb);
        }
        have_mask = fast_on;
        velsign = side * __dsub_rn(vj, __dmul_rn(alpha2, sjj));  // a.f (:110)
        if (velsign < 0.0) break;                                // :112
        LQG_TICK(9);                                             // 9: rotate + reflect
      }
      if (velsign < 0.0) {                                       // :117
        ++n_vel;
        if (++rcount > p.max_restarts) status = LQG_ERR_RESTART_CAP;
        continue;
      }
      // ---- end point: bb = sin(tt) a + cos(tt) b (:119), tmg's check min(F bb + g) >= 0 (:121)
      //      in its eta form, and the exact check of the value that would be emitted.  After fast
      //      bounces (S, C) are the angle-subtracted values (at most DCAP steps since the last exact
      //      re-basing: ~1e-14); otherwise tt is exact and they are recomputed from it ----
      if (n_dl == 0) sincos(tt, &S, &C);
      bool fine = true;
      if constexpr (!RS) {
        // One pass: end point, both checks, and the sample itself.  The end point is parked in the x_a copy
        // (x_a is consumed; a restart overwrites it with the next momentum and keeps x_b at the wall, as
        // tmg does), the sample is written before the verdict is known — a rejected end point is simply
        // overwritten when the transition is redone — and an accepted one becomes x_b by swapping the
        // two copies.
        const bool emit = s >= p.burn_in;                        // tmg:R/tmg.R:105 drops burn-in
        double* dst = p.out + ((size_t)chain * p.n_per_chain + (emit ? s - p.burn_in : 0)) * d;
#pragma unroll (CPT < 8 ? CPT : 8)
        for (int k = 0; k < CPT; ++k) {
          const int i = tid + k * T;
          const double bb = fma(xa_s[i], S, xb_s[i] * C);        // lastSample = bb (:119, :123)
          xa_s[i] = bb;
          if (i < d) {
            // (loads first, bitwise &: a short-circuit && would chain the L2 latencies of mu, lb, ub one after another)
            const double mu_i = __ldg(p.mu + i), lb_i = __ldg(p.lb + i), ub_i = __ldg(p.ub + i);
            const double e = mu_i + bb;                          // tmg:R/tmg.R:106
            const double gl = GS ? glo_s[i] : __ldg(p.glo + i), gh = GS ? ghi_s[i] : __ldg(p.ghi + i);
            fine &= (bb + gl >= 0.0) & (gh - bb >= 0.0) & (e >= lb_i) & (e <= ub_i);
            if (emit) __stcs(dst + i, e);                        // written once (again if redone), read by later stages
          }
        }
        fine = (G > 1) ? group_and<NCH, T>(gr.id, fine) : __all_sync(0xffffffffu, fine);
        if (!fine) {                                             // :130
          ++n_chk;
          if (++rcount > p.max_restarts) status = LQG_ERR_RESTART_CAP;
          continue;
        }
        double* t_ = xa_s; xa_s = xb_s; xb_s = t_;
        sh.xa_s = xa_s; sh.xb_s = xb_s;
        rcount = 0;
        ++n_trans;
      } else {
  #pragma unroll
        for (int k = 0; k < CPT; ++k) {
          const int i = tid + k * T;
          if (i < d) {
            const double bb = fma(XA(k), S, XB(k) * C);
            const double e = __ldg(p.mu + i) + bb;
            const double gl = GS ? glo_s[i] : __ldg(p.glo + i), gh = GS ? ghi_s[i] : __ldg(p.ghi + i);
            fine = fine && (bb + gl >= 0.0) && (gh - bb >= 0.0) && (e >= __ldg(p.lb + i)) && (e <= __ldg(p.ub + i));
          }
        }
        fine = (G > 1) ? group_and<NCH, T>(gr.id, fine) : __all_sync(0xffffffffu, fine);
        if (!fine) {                                               // :130
          ++n_chk;
          if (++rcount > p.max_restarts) status = LQG_ERR_RESTART_CAP;
          continue;
        }
  #pragma unroll
        for (int k = 0; k < CPT; ++k) {
          const double nb = fma(XA(k), S, XB(k) * C);              // lastSample = bb (:123)
          if (RS) xb[RS ? k : 0] = nb;
          xb_s[tid + k * T] = nb;
        }
        rcount = 0;
        ++n_trans;
        if (s >= p.burn_in) {                                      // tmg:R/tmg.R:105 drops burn-in
          double* dst = p.out + ((size_t)chain * p.n_per_chain + (s - p.burn_in)) * d;
  #pragma unroll
          for (int k = 0; k < CPT; ++k) {
            const int i = tid + k * T;
            if (i < d) __stcs(dst + i, __ldg(p.mu + i) + XB(k));   // tmg:R/tmg.R:106; written once, read by later stages
          }
        }
      }
      ++s;
      LQG_TICK(10);                                              // 10: end of the transition (check, emit)
    }
#ifdef LQG_BOX_TIMING
    if (tid == 0) for (int q_ = 0; q_ < 16; ++q_) atomicAdd(&g_box_phase[q_], ph_acc[q_]);
#endif

    // ---- persist chain state and counters ----
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
      const int i = tid + k * T;
      if (i < d) p.state[(size_t)chain * d + i] = XB(k);
    }
    if (lane == 0) atomicAdd(p.stats + 5, (unsigned long long)n_exact);
    if (tid == 0) {
      p.s_done[chain] = s; p.pos[chain] = draw; p.rcount[chain] = rcount;
      p.chain_status[chain] = status;
      atomicAdd(p.stats + 0, (unsigned long long)n_trans);
      atomicAdd(p.stats + 1, (unsigned long long)n_bounce);
      atomicAdd(p.stats + 2, (unsigned long long)n_search);
      atomicAdd(p.stats + 3, (unsigned long long)n_vel);
      atomicAdd(p.stats + 4, (unsigned long long)n_chk);
      atomicAdd(p.stats + 10, (unsigned long long)n_slow);
      atomicAdd(p.stats + 11, (unsigned long long)n_widen);
    }
    gsync();                                             // the state copy is rewritten by the next chain
  }
}

// glo = mu - lb, ghi = ub - mu (kBigG for infinite bounds), sdiag = diag(Sigma)
__global__ void box_prep_kernel(int d, const double* mu, const double* lb, const double* ub,
                                const double* Sigma, double* glo, double* ghi, double* sdiag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d) return;
  const double l = lb[i], u = ub[i], m = mu[i];
  glo[i] = isinf(l) ? kBigG : m - l;       // g2 = f mu + g with g = -lb  (tmg:R/tmg.R:51)
  ghi[i] = isinf(u) ? kBigG : u - m;       //                      g = +ub
  sdiag[i] = Sigma[(size_t)i * d + i];
}

// state[:, chain] = init[:, chain] - mu   (x_b = Ri z0 = init - mu, tmg:R/tmg.R:28)
__global__ void box_init_state_kernel(int d, int n_chains, const double* init, int64_t init_ld,
                                      const double* mu, double* state) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)d * n_chains) return;
  const int i = (int)(idx % d);
  const size_t ch = idx / d;
  state[idx] = init[ch * (size_t)init_ld + i] - mu[i];
}

__global__ void box_fork_kernel(int d, int n_chains, int F, int64_t g0, int64_t root0, const double* __restrict__ root_state,
                                const int64_t* __restrict__ root_pos, double* __restrict__ state, int64_t* __restrict__ pos,
                                int* __restrict__ s_done, int s0) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)d * n_chains) return;
  const int i = (int)(idx % d);
  const int c = (int)(idx / d);
  const int64_t g = g0 + c;
  const int64_t rho = (g - root0) / F;
  state[idx] = root_state[(size_t)rho * d + i];
  if (i == 0) {
    pos[c] = (g % F == 0) ? root_pos[rho] : 0;
    s_done[c] = s0;
  }
}

// chains that still have transitions to run (and did not fail) -> next round's work list
__global__ void box_compact_kernel(int n_chains, int total, const int* s_done, const int* status,
                                   int* active_out, int* n_active_out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_chains) return;
  if (s_done[c] < total && status[c] == 0) active_out[atomicAdd(n_active_out, 1)] = c;
  if (status[c] == 0) atomicMin(n_active_out + 1, s_done[c]);
}

// Plans the next round (one CTA) and publishes the round counters into mapped pinned host memory with
// plain stores, so the host can read them after a stream sync without a D2H copy (which would queue
// behind the bulk sample copies on the copy engine).
//   cap   = momenta per chain of the round, from the slowest chain: the whole pool while it has more than
//           a pool's worth of transitions left, then what it still needs (x1.25 + 4 for restarts); with a
//           consumer that works round by round (short_tail) the last stretch is cut into halving rounds,
//           so that the part no compute can overlap (the final round's samples) is a few columns per chain
//   count = min(cap, transitions left + restart allowance) per chain: chains ahead of the slowest one
//           do not get columns they cannot use
__global__ void __launch_bounds__(1024) box_plan_kernel(int total, int max_attempts, int short_tail, const int* s_done,
                                                        const int* active, const int* ctl, int* col_off,
                                                        volatile int* host_mapped) {
  __shared__ int wt[32];
  __shared__ int carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_active = ctl[0];
  const int s_min = min(ctl[1], total);
  const int rem_min = total - s_min;
  int cap = max(4, min(max_attempts, (int)(rem_min * 1.25) + 4));
  if (short_tail && rem_min <= 2 * max_attempts) cap = max(6, min(cap, (int)(rem_min * 0.55) + 2));
  cap = min(cap, max_attempts);                              // never past the pool allocation
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n_active; base += 1024) {
    const int a = base + threadIdx.x;
    int v = 0;
    if (a < n_active) {
      const int rem = total - s_done[active[a]];
      // restart allowance: a straggler that runs out of momenta costs a whole extra round (~1 ms of serial
      // transitions at d = 2000), a spare column per chain ~0.03 ms of GEMM: 6 + rem/6 makes that round rare
      v = min(cap, rem + 6 + rem / 6);
    }
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) wt[warp] = inc;
    __syncthreads();
    int wb = 0, tot = 0;
    for (int w = 0; w < 32; ++w) { if (w < warp) wb += wt[w]; tot += wt[w]; }
    const int c = carry;
    if (a < n_active) col_off[a] = c + wb + inc - v;
    __syncthreads();
    if (threadIdx.x == 0) carry = c + tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    col_off[n_active] = carry;
    host_mapped[0] = n_active;
    host_mapped[1] = ctl[1];
    host_mapped[2] = carry;
    host_mapped[3] = cap;
    __threadfence_system();
  }
}

// independent audit of the emitted samples: counts entries outside [lb, ub]
__global__ void box_violation_kernel(int d, size_t n_cols, const double* out, const double* lb,
                                     const double* ub, unsigned long long* count) {
  unsigned long long bad = 0;
  const size_t total = (size_t)d * n_cols;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx % d);
    const double v = out[idx];
    if (!(v >= lb[i] && v <= ub[i])) ++bad;     // NaN counts as a violation
  }
  bad = (unsigned long long)warp_sum((double)bad);
  if ((threadIdx.x & 31) == 0 && bad) atomicAdd(count, bad);
}

// grid == -1: query only, *slots_out = resident CTAs on the device
template <int G, int CPT, int THR, bool GS, int MINB, bool PACK = false, bool RS = true, int NCH = 1>
static cudaError_t launch_box_t(const BoxParams& p, int grid, cudaStream_t st, int* slots_out = nullptr) {
  constexpr size_t NS = (size_t)CPT * 32 * G;
  const size_t smem = NS * (2 * NCH + (GS ? 2 : 0) + (THR == 1 ? 2 : 0)) * sizeof(double);
  auto kern = hmc_box_kernel<G, CPT, THR, GS, MINB, PACK, RS, NCH>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 0, dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * G * NCH, smem);
  if (slots_out) {
    *slots_out = sms * (per_sm > 0 ? per_sm : 1) * NCH;      // chains in flight
    return e;
  }
  if (e != cudaSuccess) return e;
  if (grid <= 0) grid = sms * (per_sm > 0 ? per_sm : 1);     // persistent: every resident slot, once
  // fewer chains than chain slots: spread them over the SMs first (a CTA's groups pull from one queue)
  if ((long long)grid * NCH > p.n_active) grid = std::max(1, std::min(grid, p.n_active));
  if (p.n_active <= 0) return cudaSuccess;
  // Sigma (one column per bounce, 32 MB at d = 2000) should stay in L2 while the momentum pool and the
  // samples stream through it: persisting access-policy window on this launch
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(32 * G * NCH); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  int n_attr = 0;
  static const size_t persist_max = [] {
    int dev_ = 0, v = 0;
    cudaGetDevice(&dev_);
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxPersistingL2CacheSize, dev_) != cudaSuccess) { cudaGetLastError(); return (size_t)0; }
    if (v > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)v) != cudaSuccess) { cudaGetLastError(); return (size_t)0; }
    return (size_t)v;
  }();
  static const char* env_l2 = getenv("LQG_L2_PERSIST");             // tuning aid: 0 disables
  const size_t sig_bytes = (size_t)p.d * p.d * sizeof(double);
  if (persist_max > 0 && sig_bytes >= ((size_t)4 << 20) && !(env_l2 && env_l2[0] == '0')) {
    int max_win = 0;
    cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, dev);
    attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
    attr[0].val.accessPolicyWindow.base_ptr = const_cast<double*>(p.Sigma);
    attr[0].val.accessPolicyWindow.num_bytes = std::min(sig_bytes, (size_t)(max_win > 0 ? max_win : 0));
    attr[0].val.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)persist_max / (double)sig_bytes);
    attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (attr[0].val.accessPolicyWindow.num_bytes > 0) n_attr = 1;
  }
  cfg.attrs = attr; cfg.numAttrs = n_attr;
  e = cudaLaunchKernelEx(&cfg, kern, p);
  ++launch_counter();
  return e != cudaSuccess ? e : cudaGetLastError();
}

// picks (warps per chain, coordinates per thread) from d
static cudaError_t dispatch_box(const BoxParams& p, int grid, cudaStream_t st, int* slots) {
  const int d = p.d;
  if (d <= 32)   return launch_box_t<1, 1, 0, true, 16>(p, grid, st, slots);
  if (d <= 64)   return launch_box_t<1, 2, 0, true, 16>(p, grid, st, slots);
  if (d <= 128)  return launch_box_t<1, 4, 0, true, 16>(p, grid, st, slots);
  if (d <= 256)  return launch_box_t<2, 4, 0, true, 8>(p, grid, st, slots);
  static const char* v = getenv("LQG_BOX_VARIANT");            // tuning aid, see below
  // From 4 warps per chain on, the candidates of all warps are evaluated together (PACK): with ~13
  // survivors per warp, 8 partly filled evaluation passes become ~4 full ones ("n" = per-warp evaluation).
  const bool pack = !(v && v[0] == 'n');
  if (d <= 512)  return pack ? launch_box_t<4, 4, 0, true, 4, true>(p, grid, st, slots) : launch_box_t<4, 4, 0, true, 4>(p, grid, st, slots);
  // d in (512, 2048]: the state lives in shared memory only (RS = false).  The registers this frees
  // remove most of the spills of the search loop: cfg4 36.0 -> 33.7 ms per step.  "r" = register state.
  const bool rs = v && v[0] == 'r';
  if (d <= 1024) {
    // several chains per CTA (see NCH): "a" = 1 warp x 12 chains, "b" = 2 x 8, "c" = 2 x 10, "e" = 4 x 6
    if (v && v[0] == 'a') return launch_box_t<1, 32, 2, true, 1, true, false, 12>(p, grid, st, slots);
    if (v && v[0] == 'b') return launch_box_t<2, 16, 1, true, 1, true, false, 8>(p, grid, st, slots);
    if (v && v[0] == 'c') return launch_box_t<2, 16, 2, true, 1, true, false, 10>(p, grid, st, slots);
    if (v && v[0] == 'e') return launch_box_t<4, 8, 1, true, 1, true, false, 6>(p, grid, st, slots);
    if (!pack) return launch_box_t<4, 8, 1, true, 4>(p, grid, st, slots);
    return rs ? launch_box_t<4, 8, 1, true, 4, true, true>(p, grid, st, slots) : launch_box_t<4, 8, 1, true, 4, true, false>(p, grid, st, slots);
  }
  if (d <= 2048) {
    // Fewer, longer chains do less burn-in work for the same number of samples: 2 CTAs/SM (128
    // registers) is the default.  Measured alternatives: "3" = 3 CTAs/SM (80 registers, thresholds
    // recomputed from the bound offsets: 38.4 ms for 444 chains, same work rate), "4" = 4 CTAs/SM
    // (64 registers, bound offsets through L1: 49.3 ms).
    // several chains per CTA (see NCH): "a" = 2 warps x 5 chains, "b" = 4 x 5, "c" = 2 x 4 (threshold tables), "e" = 4 x 4, "f" = 4 x 3
    if (v && v[0] == 'a') return launch_box_t<2, 32, 2, true, 1, true, false, 5>(p, grid, st, slots);
    if (v && v[0] == 'b') return launch_box_t<4, 16, 2, true, 1, true, false, 5>(p, grid, st, slots);
    if (v && v[0] == 'c') return launch_box_t<2, 32, 1, true, 1, true, false, 4>(p, grid, st, slots);
    if (v && v[0] == 'e') return launch_box_t<4, 16, 1, true, 1, true, false, 4>(p, grid, st, slots);
    if (v && v[0] == '3') return launch_box_t<8, 8, 2, true, 3, true, false>(p, grid, st, slots);
    if (v && v[0] == '4') return launch_box_t<8, 8, 2, false, 4, true, false>(p, grid, st, slots);
    if (!pack) return launch_box_t<8, 8, 1, true, 2>(p, grid, st, slots);
    if (rs) return launch_box_t<8, 8, 1, true, 2, true, true>(p, grid, st, slots);
    if (v && v[0] == 'g') return launch_box_t<8, 8, 1, true, 2, true, false>(p, grid, st, slots);   // one chain per CTA of 8 warps, 2 CTAs/SM (296 chains)
    // default: 3 chains per CTA, 4 warps each (444 chains in flight): 27 % more hit searches per second
    // than 2 x 8 warps; with 100 burn-in transitions per chain and 125 000 samples per GPU that is
    // 57.2 vs 58.5 ms per step (4 and 5 chains per CTA search faster still, but their burn-in costs more)
    return launch_box_t<4, 16, 1, true, 1, true, false, 3>(p, grid, st, slots);
  }
  if (d <= 4096) return pack ? launch_box_t<8, 16, 1, true, 1, true>(p, grid, st, slots) : launch_box_t<8, 16, 1, true, 1>(p, grid, st, slots);
  if (d <= 8192) return pack ? launch_box_t<16, 16, 2, false, 1, true>(p, grid, st, slots) : launch_box_t<16, 16, 2, false, 1>(p, grid, st, slots);
  return cudaErrorInvalidValue;
}
// parity tooling: the two device hit-time functions over arrays of (f.a, f.b, g)
__global__ void hit_time_batch_kernel(size_t n, const double* fa, const double* fb, const double* g, double klim,
                                      double* t_exact, int* cls, double* key, double* st, double* ct) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    if (t_exact) t_exact[i] = hit_time_exact(fa[i], fb[i], g[i]);
    if (cls) {
      FastHit h;
      h.key = 0.0; h.st = 0.0; h.ct = 0.0;
      const int c = hit_fast(fa[i], fb[i], g[i], klim, h);
      cls[i] = c;
      if (key) key[i] = c == kFastHit ? h.key : 0.0;
      if (st) st[i] = c == kFastHit ? h.st : 0.0;
      if (ct) ct[i] = c == kFastHit ? h.ct : 0.0;
    }
  }
}
cudaError_t launch_hit_time_batch(size_t n, const double* fa, const double* fb, const double* g, double klim,
                                  double* t_exact, int* cls, double* key, double* st, double* ct, cudaStream_t stream) {
  hit_time_batch_kernel<<<148 * 8, 256, 0, stream>>>(n, fa, fb, g, klim, t_exact, cls, key, st, ct);
  ++launch_counter();
  return cudaGetLastError();
}

// diagnostic build: reads (and clears) the phase table; a normal build reports zeros
cudaError_t box_phase_cycles(unsigned long long* out16) {
#ifdef LQG_BOX_TIMING
  cudaError_t e = cudaMemcpyFromSymbol(out16, g_box_phase, 16 * sizeof(unsigned long long));
  if (e != cudaSuccess) return e;
  unsigned long long z[16] = {0};
  return cudaMemcpyToSymbol(g_box_phase, z, sizeof(z));
#else
  for (int i = 0; i < 16; ++i) out16[i] = 0;
  return cudaSuccess;
#endif
}
cudaError_t launch_hmc_box(const BoxParams& p, int grid, cudaStream_t st) { return dispatch_box(p, grid, st, nullptr); }
int hmc_box_slots(int d) {
  BoxParams p;
  memset(&p, 0, sizeof(p));
  p.d = d;
  int slots = 0;
  if (dispatch_box(p, -1, nullptr, &slots) != cudaSuccess) { cudaGetLastError(); return 0; }
  return slots;
}

cudaError_t launch_box_compact(int n_chains, int total, const int* s_done, const int* status,
                               int* active_out, int* n_active_out, int* host_mapped, int max_attempts,
                               int short_tail, int* col_off, cudaStream_t st) {
  box_compact_kernel<<<(n_chains + 255) / 256, 256, 0, st>>>(n_chains, total, s_done, status, active_out, n_active_out);
  box_plan_kernel<<<1, 1024, 0, st>>>(total, max_attempts, short_tail, s_done, active_out, n_active_out, col_off, host_mapped);
  launch_counter() += 2;
  return cudaGetLastError();
}

cudaError_t launch_box_prep(int d, const double* mu, const double* lb, const double* ub,
                            const double* Sigma, double* glo, double* ghi, double* sdiag,
                            cudaStream_t st) {
  box_prep_kernel<<<(d + 127) / 128, 128, 0, st>>>(d, mu, lb, ub, Sigma, glo, ghi, sdiag);
  ++launch_counter();
  return cudaGetLastError();
}

cudaError_t launch_box_init_state(int d, int n_chains, const double* init, int64_t init_ld,
                                  const double* mu, double* state, cudaStream_t st) {
  const size_t total = (size_t)d * n_chains;
  box_init_state_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d, n_chains, init, init_ld, mu, state);
  ++launch_counter();
  return cudaGetLastError();
}

cudaError_t launch_box_fork(int d, int n_chains, int F, int64_t g0, int64_t root0, const double* root_state,
                            const int64_t* root_pos, double* state, int64_t* pos, int* s_done, int s0, cudaStream_t st) {
  const size_t total = (size_t)d * n_chains;
  box_fork_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d, n_chains, F, g0, root0, root_state, root_pos, state, pos, s_done, s0);
  ++launch_counter();
  return cudaGetLastError();
}

cudaError_t launch_box_violations(int d, size_t n_cols, const double* out, const double* lb,
                                  const double* ub, unsigned long long* count, cudaStream_t st) {
  box_violation_kernel<<<148 * 8, 256, 0, st>>>(d, n_cols, out, lb, ub, count);
  ++launch_counter();
  return cudaGetLastError();
}

}  // namespace lqg
