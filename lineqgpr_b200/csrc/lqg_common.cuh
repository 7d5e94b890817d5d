// Shared device/host helpers for the lineqgpr_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/lineqgpr_b200.h"

namespace lqg {

constexpr double kMinT   = 0.00001;                 // tmg:src/HmcSampler.cpp:21
constexpr double kPi     = 3.14159265358979323846;  // M_PI
constexpr double kTwoPi  = 2 * kPi;                 // 2*M_PI as the reference writes it
constexpr double kHalfPi = kPi / 2;                 // T = M_PI/2, tmg:src/HmcSampler.cpp:47
constexpr double kBigG   = 1.0e300;                 // stands in for a dropped (infinite) bound
constexpr double kFilterRel = 1.0e-9;               // safety margin of the hit-time filter

// ---------------------------------------------------------------------------------------
// error plumbing (host)
void set_error(const char* fmt, ...);
int64_t& launch_counter();
#define LQG_CUDA_TRY(expr)                                                              \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      lqg::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,               \
                     cudaGetErrorString(_e));                                           \
      return LQG_ERR_CUDA;                                                              \
    }                                                                                   \
  } while (0)

// ---------------------------------------------------------------------------------------
// Philox-4x32-10 (Salmon et al. 2011), counter-based: any draw is addressable, so the
// pre-pass GEMM and the in-kernel restart path generate the same N(0,1) for the same
// (chain, transition, restart, coordinate).
struct Philox {
  static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  static constexpr uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  __host__ __device__ static inline void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#ifdef __CUDA_ARCH__
    uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
    uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
#else
    uint64_t p0 = (uint64_t)M0 * c[0], p1 = (uint64_t)M1 * c[2];
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
    uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
  __host__ __device__ static inline void gen(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                             uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
    uint32_t c[4] = {c0, c1, c2, c3};
#pragma unroll
    for (int i = 0; i < 10; ++i) { round(c, k0, k1); k0 += W0; k1 += W1; }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
  }
};

// 53-bit uniform strictly inside (0,1)
__host__ __device__ inline double u01(uint32_t hi, uint32_t lo) {
  uint64_t b = (((uint64_t)hi << 32) | lo) >> 11;                  // 53 bits
  return ((double)b + 0.5) * (1.0 / 9007199254740992.0);
}

// Two N(0,1) for the coordinate pair (2p, 2p+1) of the `draw`-th momentum of `chain`
// (draws are consumed sequentially, d normals per (re)start, like nd(eng1) at
// tmg:src/HmcSampler.cpp:54-55).  Box-Muller in FP64.
__device__ inline void normal_pair(uint64_t seed, uint64_t chain, uint64_t draw, uint32_t pair,
                                   double& n0, double& n1) {
  uint32_t r[4];
  Philox::gen(pair, (uint32_t)draw, (uint32_t)(draw >> 32), (uint32_t)chain,
              (uint32_t)seed, (uint32_t)(seed >> 32) ^ (uint32_t)(chain >> 32) * 0x85EBCA6Bu, r);
  double u1 = u01(r[0], r[1]), u2 = u01(r[2], r[3]);
  double rad = sqrt(-2.0 * log(u1));
  double s, c;
  sincospi(2.0 * u2, &s, &c);
  n0 = rad * c; n1 = rad * s;
}

// ---------------------------------------------------------------------------------------
// One constraint's hit time, tmg:src/HmcSampler.cpp:157-181.  fa = f.a, fb = f.b.  Returns the
// candidate t, or 0 when the constraint offers no hit (u <= |g|, or t <= min_t).  The grazing /
// dead-zone branch follows the reference operation for operation (no FMA contraction there on
// purpose: the CPU reference is compiled without it and these few flops decide discrete events);
// the well-conditioned branch takes t1 from ONE atan2 of the complex product, which equals tmg's
// acos - atan2 value to a few ulp, not bit for bit — tests/test_gpu_parity_full.py bounds the
// difference on 10^7 random triples (verdict flips only within 1e-11/sin(theta0) of a threshold).
// Deliberately NOT inlined: it runs for the few survivors of the filter only, and keeping its
// atan2/acos/sqrt temporaries out of the callers' register budget buys occupancy.
static __device__ __noinline__ double hit_time_exact(double fa, double fb, double g) {
  const double u2 = __dadd_rn(__dmul_rn(fa, fa), __dmul_rn(fb, fb));
  double u = sqrt(u2);                                                  // :159
  if (!(u > g && u > -g)) return 0.0;                                   // :160
  double phi, t1;
  const double w2 = fma(-g, g, u2);                                     // u^2 - g^2 = (u sin theta0)^2
  if (w2 > 1e-4 * u2) {
    // Well-conditioned wall (not grazing): t1 = acos(-g/u) - atan2(-fa, fb) is the angle of
    // (-g + i w)(fb + i fa), w = sqrt(u^2 - g^2): one atan2 instead of atan2 + acos + a division,
    // equal to tmg's value to a few ulp.  t2 (below) needs phi only as -t1 - 2 phi; with
    // theta0 = t1 + phi it is written t2 = t1 - 2 theta0' where theta0' is recovered from the
    // same complex form, so phi is computed lazily only if the dead zones require t2.
    const double w = sqrt(w2);
    t1 = atan2(fma(w, fb, -g * fa), fma(-g, fb, -w * fa));
    if (t1 < 0) t1 += kTwoPi;
    if (fb + g > 1e-9 * (fabs(fb) + fabs(g)) && t1 > 2.0 * kMinT && t1 < kTwoPi - 2.0 * kMinT)
      return t1;           // strictly feasible start: t1 < t2 always, no dead zone
    phi = atan2(-fa, fb);
  } else {
    phi = atan2(-fa, fb);                                               // :161
    t1 = acos(-g / u) - phi;                                            // :162
    if (t1 < 0) t1 += kTwoPi;                                           // :165
  }
  if (fabs(t1) < kMinT) t1 = 0;                                         // :166
  else if (fabs(t1 - kTwoPi) < kMinT) t1 = 0;                           // :167
  double t2 = __dsub_rn(-t1, __dmul_rn(2.0, phi));                      // :170
  if (t2 < 0) t2 += kTwoPi;                                             // :171
  if (t2 < 0) t2 += kTwoPi;                                             // :172
  if (fabs(t2) < kMinT) t2 = 0;                                         // :174
  else if (fabs(t2 - kTwoPi) < kMinT) t2 = 0;                           // :175
  double t = t1;                                                        // :178
  if (t1 == 0) t = t2;                                                  // :179
  else if (t2 == 0) t = t1;                                             // :180
  else t = (t1 < t2 ? t1 : t2);                                         // :181
  return t > kMinT ? t : 0.0;                                           // :183 (first half)
}

// ---------------------------------------------------------------------------------------
// The same hit time without inverse trigonometry.  With u = |(fa, fb)|, w = sqrt(u^2 - g^2) the two
// zero crossings of y(s) = fa sin s + fb cos s + g are the angles of
//     exit   (t1 of tmg):  (-g + i w)(fb + i fa) / u^2
//     entry  (t2 of tmg):  (-g - i w)(fb + i fa) / u^2
// (tmg:src/HmcSampler.cpp:161-162,170 written as complex products), so cos t and sin t of either are
// four products and one square root away, and for t in (0, pi) the value tan(t/2) = sin t / (1 + cos t)
// orders like t with uniform resolution.  A sampleNext step only ever needs: WHICH wall is hit first,
// whether that is before the remaining time tt, and sin t / cos t of that hit for the rotation
// (:82, :91-92).  hit_fast returns, for one wall, the key tan(t/2) of tmg's hit time t with its cos and
// sin (all scaled by u^2 = `u2`), or says that the wall cannot be hit within the remaining time
// (key > klim), or that the case is too close to one of tmg's discrete thresholds to be decided
// without its own arithmetic: u ~ |g| (:160), t1 or t2 within 1e-6 relative of the 1e-5 dead zone
// (:166-167, :174-175).  Callers resolve those (and near-ties between walls, and hits within 1e-9 of
// tt) with hit_time_exact, so every discrete event is still decided by tmg's formula; the exact t of
// the walls that were hit is evaluated later, off the critical path, to keep tt = tt - t (:90) exact.
constexpr int kFastNone = 0, kFastHit = 1, kFastSlow = 2;
constexpr double kSinMinT = 9.9999999998333333e-06;                     // sin(min_t)
struct FastHit { double key, st, ct; };                                 // tan(t/2), sin t, cos t
__device__ __forceinline__ int hit_fast(double fa, double fb, double g, double klim, FastHit& h) {
  const double u2 = fma(fa, fa, fb * fb);
  const double w2 = fma(-g, g, u2);
  if (!(w2 > 1e-6 * u2)) return (w2 < -1e-6 * u2) ? kFastNone : kFastSlow;   // out of reach (:160) | grazing
  const double w = sqrt(w2);
  const double ru = 1.0 / u2;                                           // independent of the sqrt: overlaps it
  const double gb = g * fb, ga = g * fa, wa = w * fa, wb = w * fb;
  const double c1 = -gb - wa, s1 = wb - ga;                             // exit crossing, t1
  double c2 = -gb + wa, s2 = -wb - ga;                                  // entry crossing, t2 = -t1 - 2 phi
  const double dz = kSinMinT * u2, band = 1e-6 * dz;
  bool amb = (c1 > 0.0) & (fabs(fabs(s1) - dz) <= band);
  const bool dead1 = (c1 > 0.0) & (fabs(s1) < dz);                      // |t1| or |t1 - 2 pi| < min_t (:166-167)
  if (dead1) {                                                          // t1 := 0  =>  t2 = -2 phi (:170, Q6)
    c2 = fma(fb, fb, -fa * fa);
    s2 = 2.0 * fa * fb;
  }
  amb |= (c2 > 0.0) & (fabs(fabs(s2) - dz) <= band);
  const bool dead2 = (c2 > 0.0) & (fabs(s2) < dz);                      // (:174-175)
  // a crossing matters when it is alive, in (0, pi) and not later than the remaining time:
  // tan(t/2) = s / (u2 + c) <= klim
  const bool ok1 = !dead1 & (s1 > 0.0) & (s1 <= klim * (u2 + c1));
  const bool ok2 = !dead2 & (s2 > 0.0) & (s2 <= klim * (u2 + c2));
  if (amb) return kFastSlow;
  if (!(ok1 | ok2)) return kFastNone;
  const bool use1 = ok1 & (!ok2 | (c1 >= c2));                          // the earlier one: larger cosine (:178-181)
  const double c = use1 ? c1 : c2, sn = use1 ? s1 : s2;
  h.key = sn / (u2 + c);
  h.st = sn * ru;
  h.ct = c * ru;
  return kFastHit;
}

// Conservative filter in front of hit_time_exact.  y(t) = fa sin t + fb cos t + g is the
// constraint value along the trajectory; u = sqrt(fa^2 + fb^2) its amplitude.
//  (1) tmg only considers a wall when u > |g| (tmg:src/HmcSampler.cpp:160): if u < g - margin
//      the wall is out of reach for this whole orbit.
//  (2) only hits with t <= tt can matter to sampleNext (:82 breaks when tt < t), and such a hit
//      needs a zero of y on [0, tt].  On an interval no longer than pi/2, y has at most one
//      interior extremum; it is a minimum iff y'(0) = fa < 0 and y'(tt) = fa cos tt - fb sin tt > 0
//      (then min y = g - u, covered by (1)); otherwise min y = min(y(0), y(tt)).
// A constraint is skipped only when those bounds clear zero by a margin far larger than any
// rounding difference, so every constraint that can win the arg-min is still evaluated by
// hit_time_exact and tie/threshold semantics are untouched.
//   S = sin tt, C = cos tt; margin ~ 1e-9 * scale of (fa, fb, g)
__device__ __forceinline__ bool may_hit(double fa, double fb, double g, double S, double C,
                                        double margin) {
  const double gm = g - margin;
  if (gm > 0.0 && gm * gm > fma(fa, fa, fb * fb)) return false;      // (1) u < g - margin
  const double y0 = fb + g;
  const double y1 = fma(fa, S, fma(fb, C, g));
  const double d1 = fma(fa, C, -fb * S);
  const bool interior_min = (fa < 0.0) && (d1 > 0.0);
  return !(y0 > margin && y1 > margin) || interior_min;               // (2)
}

// (t, key) lexicographic min with t == 0 meaning "none": lowest key wins ties, which is the
// reference's "first index wins" (strict < at tmg:src/HmcSampler.cpp:183).
struct Hit {
  double t; int key;
};
__device__ __forceinline__ bool hit_better(double t, int key, double bt, int bkey) {
  return t > 0.0 && (bt == 0.0 || t < bt || (t == bt && key < bkey));
}
__device__ __forceinline__ Hit warp_min_hit(Hit h) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    double ot = __shfl_xor_sync(0xffffffffu, h.t, off);
    int ok = __shfl_xor_sync(0xffffffffu, h.key, off);
    if (hit_better(ot, ok, h.t, h.key)) { h.t = ot; h.key = ok; }
  }
  return h;
}
// Arg-min of (t, key) over a warp with the hardware warp-reduce (REDUX): a positive double
// orders like its bit pattern, "no hit" is +inf, ties on t go to the lowest key (tmg's strict <
// at tmg:src/HmcSampler.cpp:183).  tb = time bits (kNoHitBits when none).
// Returns the winning lane (every lane gets the same tb_min / key_min).
constexpr unsigned long long kNoHitBits = 0x7ff0000000000000ull;      // bits of +inf
__device__ __forceinline__ unsigned long long hit_bits(double t) {
  return t > 0.0 ? (unsigned long long)__double_as_longlong(t) : kNoHitBits;
}
__device__ __forceinline__ double hit_from_bits(unsigned long long tb) {
  return tb == kNoHitBits ? 0.0 : __longlong_as_double((long long)tb);
}
__device__ __forceinline__ int warp_argmin_hit(unsigned long long tb, int key,
                                               unsigned long long& tb_min, int& key_min) {
  const unsigned hi = (unsigned)(tb >> 32), lo = (unsigned)tb;
  const unsigned hi_min = __reduce_min_sync(0xffffffffu, hi);
  const unsigned lo_min = __reduce_min_sync(0xffffffffu, hi == hi_min ? lo : 0xffffffffu);
  const bool is_min = (hi == hi_min) && (lo == lo_min);
  const unsigned k_min = __reduce_min_sync(0xffffffffu, is_min ? (unsigned)key : 0x7fffffffu);
  tb_min = ((unsigned long long)hi_min << 32) | lo_min;
  key_min = (int)k_min;
  const unsigned owner = __ballot_sync(0xffffffffu, is_min && (unsigned)key == k_min);
  return __ffs(owner) - 1;
}

// ---------------------------------------------------------------------------------------
// mbarrier + bulk-async copy (TMA engine, SASS UBLKCP) helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------------------------------
// log of the N(0,1) mass of [a, b], stable in both tails (ExpT: proposals and set-up)
constexpr double kInvSqrt2 = 0.7071067811865476;
// log of the upper tail of N(0,1), stable for large a:  log(erfcx(a/sqrt2)/2) - a^2/2
__device__ __forceinline__ double log_upper(double a) {
  if (a > 0.0) return log(0.5 * erfcx(a * kInvSqrt2)) - 0.5 * a * a;
  return log(0.5 * erfc(a * kInvSqrt2));
}
static __device__ double lnNpr(double a, double b) {
  if (a > 0.0) {                                   // 0 < a < b
    const double pa = log_upper(a), pb = log_upper(b);
    return pa + log1p(-exp(pb - pa));
  }
  if (b < 0.0) {                                   // a < b < 0
    const double pa = log_upper(-a), pb = log_upper(-b);
    return pb + log1p(-exp(pa - pb));
  }
  return log1p(-0.5 * erfc(-a * kInvSqrt2) - 0.5 * erfc(b * kInvSqrt2));   // a <= 0 <= b
}

// Several chains may share one CTA (NCH > 1): a chain is then run by a GROUP of G warps with its own
// named barrier (id 1 + group), its own state copy, queues and slots; what the groups share is the read-only
// per-coordinate tables in shared memory.  With NCH == 1 the group is the CTA and the barrier is barrier 0.
struct Grp {
  int id;      // group (chain slot) inside the CTA
  int tid;     // thread index inside the group
};
template <int NCH, int T>
__device__ __forceinline__ void group_sync(int grp) {
  if (NCH == 1) __syncthreads();
  else asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "n"(T) : "memory");
}
template <int NCH, int T>
__device__ __forceinline__ bool group_and(int grp, bool v) {
  if (NCH == 1) return __syncthreads_and(v) != 0;
  unsigned r;
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "setp.ne.u32 p, %1, 0;\n"
      "bar.red.and.pred q, %2, %3, p;\n"
      "selp.u32 %0, 1, 0, q;\n"
      "}\n"
      : "=r"(r)
      : "r"((unsigned)v), "r"(grp + 1), "n"(T)
      : "memory");
  return r != 0;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

}  // namespace lqg
