// Summary bands of the simulated paths: row means and R type-7 quantiles of ysim (n_t x N),
//   qtls  <- apply(ysim, 1, quantile, probs = probs)      R/lineqGPutils.R:422, :296
//   ymean <- apply(ysim, 1, mean)                          R/lineqGPutils.R:423, :294
//
// ysim is column-major with one COLUMN per sample, so the N samples of one test point are
// strided.  The result is EXACT (no approximation of the order statistics) on both paths:
//
// Bracketed path (default) — one coalesced pass over ysim, HBM-bound:
//   1. sample   2048 strided samples per test point are sorted in shared memory; for every
//               probability two of them bracket the wanted order statistics with a 4-sigma margin
//   2. collect  CTAs own 32 consecutive test points (lane = test point, so every load is a 256-byte
//               row segment of one sample) x a quarter of the samples: running sums for the mean,
//               the count of values below each bracket, and the values inside it (~4-10 %)
//               appended to a per-(test point, probability) buffer
//   3. select   one CTA per test point: rank inside the bracket = rank - count below; radix select
//               on the small buffer (L2 resident), type-7 interpolation
//   A bracket that misses (probability ~6e-5 per target), overflows (heavy ties) or a tiny N
//   sends the test point to the exact path below; more than 64 such rows send the whole slab.
//
// Exact path (fallback) — tiled transpose, then one CTA per test point: deterministic tree sum
// and a most-significant-digit radix select (8 passes of 8 bits over an order-preserving 64-bit
// key, per-target histograms in shared memory, warp-aggregated atomics).
#include <string.h>

#include <vector>

#include "lqg_common.cuh"
#include "lqg_kernels.h"

namespace lqg {

namespace {

constexpr int MAXT = 16;            // order statistics per row (2 per probability)
constexpr int BT = 256;
constexpr int UNR = 8;               // loads in flight per thread while streaming a row

__device__ __forceinline__ uint64_t ordered_key(double x) {
  uint64_t b = (uint64_t)__double_as_longlong(x);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_to_double(uint64_t k) {
  uint64_t b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

struct BandTargets {
  int n;                       // number of order statistics (2 * nprobs)
  long long rank[MAXT];        // 0-based ranks
  double h[MAXT / 2];          // interpolation weights per probability
};

}  // namespace

// out (cols x rows) = in (rows x cols)', 32x32 tiles through padded shared memory
__global__ void transpose_kernel(int rows, int64_t cols, const double* __restrict__ in, int64_t ld_in,
                                 double* __restrict__ out, int64_t ld_out) {
  __shared__ double tile[32][33];
  const int64_t c0 = (int64_t)blockIdx.x * 32;
  const int r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + threadIdx.x; const int64_t c = c0 + j;
    tile[j][threadIdx.x] = (r < rows && c < cols) ? in[(size_t)c * ld_in + r] : 0.0;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int64_t c = c0 + threadIdx.x; const int r = r0 + j;
    if (r < rows && c < cols) out[(size_t)r * ld_out + c] = tile[threadIdx.x][j];
  }
}

__global__ void __launch_bounds__(BT)
bands_kernel(int n_t, int64_t N, const double* __restrict__ yT, int64_t ldT, BandTargets tg,
             double* __restrict__ mean, double* __restrict__ qtls) {
  __shared__ unsigned int hist[MAXT][256];
  __shared__ double red[BT / 32];
  __shared__ uint64_t prefix[MAXT];
  __shared__ long long rem[MAXT];
  __shared__ int group_of[MAXT];
  __shared__ uint64_t group_prefix[MAXT];
  __shared__ int n_groups;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int row = blockIdx.x; row < n_t; row += gridDim.x) {
    const double* y = yT + (size_t)row * ldT;
    // ---- mean: per-thread strided partial sums, then a fixed-order tree (deterministic) ----
    double s = 0.0;
    {
      // 8 independent loads per thread in flight (the row is streamed from HBM/L2: latency bound otherwise)
      double part[UNR];
#pragma unroll
      for (int q = 0; q < UNR; ++q) part[q] = 0.0;
      for (int64_t j0 = 0; j0 < N; j0 += (int64_t)BT * UNR) {
#pragma unroll
        for (int q = 0; q < UNR; ++q) {
          const int64_t j = j0 + (int64_t)q * BT + tid;
          part[q] += j < N ? y[j] : 0.0;
        }
      }
#pragma unroll
      for (int q = 0; q < UNR; ++q) s += part[q];
    }
    s = warp_sum(s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < BT / 32; ++w) t += red[w];
      mean[row] = t / (double)N;
    }
    if (tid < tg.n) { prefix[tid] = 0; rem[tid] = tg.rank[tid]; }
    __syncthreads();

    // ---- MSD radix select for all targets at once ----
    for (int pass = 0; pass < 8; ++pass) {
      const int shift = 56 - 8 * pass;
      if (tid == 0) {                       // targets sharing a prefix share a histogram
        int ng = 0;
        for (int k = 0; k < tg.n; ++k) {
          int g = -1;
          for (int q = 0; q < ng; ++q) if (group_prefix[q] == prefix[k]) { g = q; break; }
          if (g < 0) { g = ng; group_prefix[ng++] = prefix[k]; }
          group_of[k] = g;
        }
        n_groups = ng;
      }
      for (int i = tid; i < MAXT * 256; i += BT) (&hist[0][0])[i] = 0u;
      __syncthreads();
      const int ng = n_groups;
      for (int64_t j0 = 0; j0 < N; j0 += (int64_t)BT * UNR) {
        double v[UNR];
#pragma unroll
        for (int q = 0; q < UNR; ++q) {                     // UNR independent loads in flight per thread
          const int64_t j = j0 + (int64_t)q * BT + tid;
          v[q] = j < N ? y[j] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < UNR; ++q) {
          const int64_t j = j0 + (int64_t)q * BT + tid;
          int g = -1; unsigned digit = 0;
          if (j < N) {
            const uint64_t key = ordered_key(v[q]);
            const uint64_t hi = pass == 0 ? 0ull : (key >> (shift + 8));
            digit = (unsigned)((key >> shift) & 0xffu);
            for (int c = 0; c < ng; ++c) if (group_prefix[c] == hi) { g = c; break; }
          }
          // warp-aggregated shared atomics: lanes with the same (group, digit) add once.  A warp in
          // which no lane matched any target prefix (the common case after the first passes) skips.
          if (__any_sync(0xffffffffu, g >= 0)) {
            const unsigned code = g < 0 ? 0xffffffffu : (unsigned)(g * 256 + (int)digit);
            const unsigned peers = __match_any_sync(0xffffffffu, code);
            if (g >= 0 && (__ffs(peers) - 1) == lane) atomicAdd(&hist[g][digit], (unsigned)__popc(peers));
          }
        }
      }
      __syncthreads();
      if (tid < tg.n) {
        const unsigned int* hh = hist[group_of[tid]];
        long long r = rem[tid], cum = 0;
        int dsel = 255;
        for (int b = 0; b < 256; ++b) {
          const long long c = hh[b];
          if (r < cum + c) { dsel = b; break; }
          cum += c;
        }
        rem[tid] = r - cum;
        prefix[tid] = (prefix[tid] << 8) | (uint64_t)dsel;
      }
      __syncthreads();
    }
    // ---- type-7 interpolation: (1-h) x[lo] + h x[hi]  (stats::quantile, R >= 4.0) ----
    if (tid < tg.n / 2) {
      const double xlo = key_to_double(prefix[2 * tid]);
      const double xhi = key_to_double(prefix[2 * tid + 1]);
      const double h = tg.h[tid];
      double qv = xlo;
      if (h == 1.0) qv = xhi;
      else if (h > 0.0 && h < 1.0 && xlo != xhi)   // no FMA contraction: R rounds both products
        qv = __dadd_rn(__dmul_rn(1.0 - h, xlo), __dmul_rn(h, xhi));
      qtls[(size_t)row * (tg.n / 2) + tid] = qv;
    }
    __syncthreads();
  }
}

// =============================================================================================
// Bracketed path
namespace {

constexpr int SAMP = 2048;           // samples per test point
constexpr int CW = 8;                // warps per collect CTA: 8 x 32 = 256 consecutive test points
constexpr int CROWS = 32 * CW;
constexpr int MAXSPLIT = 256;        // CTAs along the sample axis per 256-row tile (chosen at launch)
constexpr int MAXP = MAXT / 2;       // probabilities
constexpr int CUNR = 16;             // sample columns in flight per thread in the collect pass

struct BandPlan {
  int nprobs;
  int sampled;                       // 0: brackets are (-inf, +inf) (small N: everything is collected)
  int lo_i[MAXP], hi_i[MAXP];        // sample indices of the bracket ends (< 0 / >= SAMP: open end)
  int splits;                        // CTAs along the sample axis
  long long cap_split;               // buffer capacity per (test point, probability, split)
};

}  // namespace

// 1. brackets[row][k] = {lo key, hi key} from a sorted strided sample of the row
__global__ void __launch_bounds__(SAMP / 2)
bands_sample_kernel(int n_t, int64_t N, const double* __restrict__ ysim, int64_t ld, BandPlan pl,
                    uint64_t* __restrict__ brackets) {
  __shared__ uint64_t key[SAMP];
  const int row = blockIdx.x, tid = threadIdx.x;
  for (int e = tid; e < SAMP; e += SAMP / 2) {
    const int64_t n = ((int64_t)e * N) / SAMP;             // N < 2^31: no overflow
    key[e] = ordered_key(ysim[(size_t)n * ld + row]);
  }
  for (int k = 2; k <= SAMP; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      __syncthreads();
      const int a = 2 * j * (tid / j) + (tid % j), b = a + j;      // one compare-exchange per thread
      const uint64_t ka = key[a], kb = key[b];
      const bool up = (a & k) == 0;
      if ((ka > kb) == up) { key[a] = kb; key[b] = ka; }
    }
  __syncthreads();
  if (tid < pl.nprobs) {
    brackets[((size_t)row * pl.nprobs + tid) * 2 + 0] = pl.lo_i[tid] < 0 ? 0ull : key[pl.lo_i[tid]];
    brackets[((size_t)row * pl.nprobs + tid) * 2 + 1] = pl.hi_i[tid] >= SAMP ? ~0ull : key[pl.hi_i[tid]];
  }
}

// 2. one coalesced pass: a CTA owns 256 consecutive test points (thread = test point, so the CTA
//    reads 2 KB contiguous of every sample column) x one slice of the samples.  Everything a
//    thread needs — brackets, counters, buffer cursors, the compensated sum — lives in its
//    registers (NP is a template parameter); there is no shared memory and no atomic.
template <int NP>
__global__ void __launch_bounds__(CROWS)
bands_collect_kernel(int n_t, int64_t N, const double* __restrict__ ysim, int64_t ld, BandPlan pl,
                     const uint64_t* __restrict__ brackets, double* __restrict__ part_sum,
                     unsigned* __restrict__ below_g, unsigned* __restrict__ cnt_g, double* __restrict__ buf) {
  const int tid = threadIdx.x;
  const int row = blockIdx.x * CROWS + tid, split = blockIdx.y;
  if (row >= n_t) return;
  constexpr int NQ = NP > 0 ? NP : 1;
  uint64_t lo[NQ], hi[NQ];
  unsigned below[NQ], cnt[NQ];
  double* seg[NQ];
  const unsigned cap = (unsigned)pl.cap_split;
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    lo[k] = 0ull; hi[k] = ~0ull;
    if (pl.sampled) { lo[k] = brackets[((size_t)row * NP + k) * 2]; hi[k] = brackets[((size_t)row * NP + k) * 2 + 1]; }
    below[k] = 0u; cnt[k] = 0u;
    seg[k] = buf + (((size_t)row * NP + k) * pl.splits + split) * (size_t)pl.cap_split;
  }
  const int64_t chunk = (N + pl.splits - 1) / pl.splits;
  const int64_t n_begin = (int64_t)split * chunk, n_end = n_begin + chunk < N ? n_begin + chunk : N;
  double sum = 0.0, comp = 0.0;        // Neumaier-compensated running sum (the pass is memory bound)
  auto consume = [&](double x) {
    const double t = __dadd_rn(sum, x);
    comp = __dadd_rn(comp, fabs(sum) >= fabs(x) ? __dadd_rn(__dsub_rn(sum, t), x) : __dadd_rn(__dsub_rn(x, t), sum));
    sum = t;
    const uint64_t key = ordered_key(x);
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      below[k] += key < lo[k] ? 1u : 0u;
      if (key >= lo[k] && key <= hi[k]) {
        if (cnt[k] < cap) seg[k][cnt[k]] = x;
        ++cnt[k];
      }
    }
  };
  const double* src = ysim + (size_t)n_begin * ld + row;
  int64_t n = n_begin;
  for (; n + CUNR <= n_end; n += CUNR, src += (size_t)CUNR * ld) {
    double v[CUNR];
#pragma unroll
    for (int q = 0; q < CUNR; ++q) v[q] = __ldg(src + (size_t)q * ld);   // CUNR sample columns in flight
#pragma unroll
    for (int q = 0; q < CUNR; ++q) consume(v[q]);
  }
  for (; n < n_end; ++n, src += ld) consume(__ldg(src));
  part_sum[((size_t)row * pl.splits + split) * 2 + 0] = sum;
  part_sum[((size_t)row * pl.splits + split) * 2 + 1] = comp;
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    below_g[((size_t)split * n_t + row) * NP + k] = below[k];
    cnt_g[((size_t)split * n_t + row) * NP + k] = cnt[k];
  }
}

// 3. one CTA per test point: mean; then per probability the rank inside the bracket
//    (rank - count below) is located by LINEAR binning of the order-preserving keys between the
//    smallest and the largest collected key — a bracket is a narrow value range, so its keys are
//    close to uniform there and one histogram of 1024 bins isolates a few dozen candidates
//    (an MSD radix pass would see the same leading digits in every element).  Passes over the
//    collected values: min/max, histogram, gather of the selected bin (repeated on the bin's key
//    range while it holds more than LIST candidates), then exact ranking of the short list in
//    shared memory.  Warps walk whole segments, lanes stride inside them.
constexpr int SEL_BINS = 1024;
constexpr int SEL_LIST = 1024;

__global__ void __launch_bounds__(BT)
bands_select_kernel(int n_t, int64_t N, BandPlan pl, BandTargets tg, const double* __restrict__ part_sum,
                    const unsigned* __restrict__ below_g, const unsigned* __restrict__ cnt_g,
                    const double* __restrict__ buf, double* __restrict__ mean, double* __restrict__ qtls,
                    int* __restrict__ fail) {
  __shared__ unsigned hist[SEL_BINS];
  __shared__ uint64_t list[SEL_LIST];
  __shared__ unsigned c[MAXSPLIT];
  __shared__ unsigned long long kmin_s, kmax_s, next_key, ans_lo, ans_hi;
  __shared__ unsigned n_list;
  __shared__ long long rem_s, below_s, n_in_s;
  __shared__ int sel_s, flag_s;
  const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int np = pl.nprobs, ns = pl.splits;
  if (tid == 0) {
    double t = 0.0, cc = 0.0;                                  // fixed order, compensated
    for (int i = 0; i < 2 * ns; ++i) {
      const double x = part_sum[(size_t)row * (2 * ns) + i], u = __dadd_rn(t, x);
      cc = __dadd_rn(cc, fabs(t) >= fabs(x) ? __dadd_rn(__dsub_rn(t, u), x) : __dadd_rn(__dsub_rn(x, u), t));
      t = u;
    }
    mean[row] = (t + cc) / (double)N;
  }
  bool failed = false;
  // applies f(key) to every collected value of (row, k): warps take segments, lanes stride inside
  auto for_each_key = [&](const double* seg0, auto&& f) {
    if (ns == 1) {                                             // one segment (fused product): the whole CTA strides over it
      const unsigned cs = c[0];
      for (unsigned j = tid; j < cs; j += BT) f(ordered_key(seg0[j]));
      return;
    }
    for (int sp = warp; sp < ns; sp += BT / 32) {
      const double* seg = seg0 + (size_t)sp * pl.cap_split;
      const unsigned cs = c[sp];
      for (unsigned j = lane; j < cs; j += 32) f(ordered_key(seg[j]));
    }
  };
  for (int k = 0; k < np; ++k) {
    __syncthreads();
    if (tid < ns) c[tid] = cnt_g[((size_t)tid * n_t + row) * np + k];
    if (tid == 0) {
      long long below = 0, n_in = 0;
      int bad = 0;
      for (int sp = 0; sp < ns; ++sp) {
        below += below_g[((size_t)sp * n_t + row) * np + k];
        const unsigned cs = cnt_g[((size_t)sp * n_t + row) * np + k];
        bad |= (long long)cs > pl.cap_split;
        n_in += cs;
      }
      below_s = below; n_in_s = n_in; flag_s = bad;
      kmin_s = ~0ull; kmax_s = 0ull; next_key = ~0ull;
    }
    __syncthreads();
    const long long r_lo = tg.rank[2 * k] - below_s, r_hi = tg.rank[2 * k + 1] - below_s;
    if (flag_s || r_lo < 0 || r_hi >= n_in_s) {               // uniform across the CTA
      failed = true;
      if (tid == 0) qtls[(size_t)row * np + k] = __longlong_as_double(0x7ff8000000000000ll);
      continue;
    }
    const double* seg0 = buf + ((size_t)row * np + k) * ns * (size_t)pl.cap_split;
    // ---- pass A: key range of the collected values ----
    {
      unsigned long long mn = ~0ull, mx = 0ull;
      for_each_key(seg0, [&](uint64_t key) { mn = key < mn ? key : mn; mx = key > mx ? key : mx; });
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long a = __shfl_xor_sync(0xffffffffu, mn, o), b = __shfl_xor_sync(0xffffffffu, mx, o);
        mn = a < mn ? a : mn; mx = b > mx ? b : mx;
      }
      if (lane == 0) { atomicMin(&kmin_s, mn); atomicMax(&kmax_s, mx); }
    }
    if (tid == 0) rem_s = r_lo;
    __syncthreads();
    uint64_t klo = kmin_s, khi = kmax_s;                        // current key range [klo, khi] holding rank rem_s
    bool resolved = false;
    // ---- narrow the range until its candidates fit the shared list ----
    for (int round = 0; round < 8; ++round) {
      const uint64_t span = khi - klo;
      // bin(key) = floor((key - klo) * SEL_BINS / (span + 1)) computed as a 128-bit product with
      // ceil(2^64 * SEL_BINS / (span + 1)) would not be monotone-safe; use the exact quotient form
      // on a down-shifted difference instead: shift so that (span >> sh) < 2^53, then double math
      // is exact and monotone.
      int sh = 0;
      while ((span >> sh) >= (1ull << 52)) ++sh;
      const double scale = (double)SEL_BINS / ((double)(span >> sh) + 1.0);
      auto bin_of = [&](uint64_t key) -> int {
        const int b = (int)((double)((key - klo) >> sh) * scale);
        return b < SEL_BINS ? b : SEL_BINS - 1;
      };
      for (int i = tid; i < SEL_BINS; i += BT) hist[i] = 0u;
      if (tid == 0) n_list = 0u;
      __syncthreads();
      for_each_key(seg0, [&](uint64_t key) { if (key >= klo && key <= khi) atomicAdd(&hist[bin_of(key)], 1u); });
      __syncthreads();
      if (tid == 0) {
        long long r = rem_s, cum = 0;
        int bsel = SEL_BINS - 1;
        for (int b = 0; b < SEL_BINS; ++b) {
          const long long cc = hist[b];
          if (r < cum + cc) { bsel = b; break; }
          cum += cc;
        }
        rem_s = r - cum;
        sel_s = bsel;
      }
      __syncthreads();
      const int bsel = sel_s;
      const unsigned in_bin = hist[bsel];
      if (in_bin <= SEL_LIST || span == 0) {
        // gather the bin (all of it fits, or every key is equal) and rank exactly
        for_each_key(seg0, [&](uint64_t key) {
          if (key >= klo && key <= khi && bin_of(key) == bsel) {
            const unsigned pos = atomicAdd(&n_list, 1u);
            if (pos < SEL_LIST) list[pos] = key;
          }
        });
        __syncthreads();
        resolved = true;
        break;
      }
      // too many candidates: restrict the key range to this bin's members and bin again
      if (tid == 0) { kmin_s = ~0ull; kmax_s = 0ull; }
      __syncthreads();
      {
        unsigned long long mn = ~0ull, mx = 0ull;
        for_each_key(seg0, [&](uint64_t key) {
          if (key >= klo && key <= khi && bin_of(key) == bsel) { mn = key < mn ? key : mn; mx = key > mx ? key : mx; }
        });
        for (int o = 16; o > 0; o >>= 1) {
          const unsigned long long a = __shfl_xor_sync(0xffffffffu, mn, o), b = __shfl_xor_sync(0xffffffffu, mx, o);
          mn = a < mn ? a : mn; mx = b > mx ? b : mx;
        }
        if (lane == 0) { atomicMin(&kmin_s, mn); atomicMax(&kmax_s, mx); }
      }
      __syncthreads();
      klo = kmin_s; khi = kmax_s;
      __syncthreads();
    }
    if (!resolved) {                                           // cannot happen for finite data; never guess
      failed = true;
      if (tid == 0) qtls[(size_t)row * np + k] = __longlong_as_double(0x7ff8000000000000ll);
      continue;
    }
    // ---- exact ranking inside the list: the element with exactly rem_s smaller-or-earlier entries ----
    const unsigned nl = n_list < SEL_LIST ? n_list : SEL_LIST;
    const long long want = rem_s;
    if (tid == 0) { ans_lo = ~0ull; ans_hi = ~0ull; }
    __syncthreads();
    const bool all_equal = (khi == klo);
    if (all_equal) {
      if (tid == 0) { ans_lo = klo; ans_hi = (want + 1 < (long long)n_list) ? klo : ~0ull; }
    } else {
      for (unsigned e = tid; e < nl; e += BT) {
        const uint64_t ke = list[e];
        long long less = 0;
        for (unsigned j = 0; j < nl; ++j) { const uint64_t kj = list[j]; less += (kj < ke) || (kj == ke && j < e); }
        if (less == want) ans_lo = ke;
        if (less == want + 1) ans_hi = ke;
      }
    }
    __syncthreads();
    const uint64_t key_lo = ans_lo;
    uint64_t key_hi = ans_hi;
    if (r_hi == r_lo) key_hi = key_lo;
    else if (key_hi == ~0ull) {
      // the next order statistic lies outside the list: the smallest collected key above key_lo
      unsigned long long nx = ~0ull;
      for_each_key(seg0, [&](uint64_t key) { if (key > key_lo && key < nx) nx = key; });
      for (int o = 16; o > 0; o >>= 1) { const unsigned long long a = __shfl_xor_sync(0xffffffffu, nx, o); nx = a < nx ? a : nx; }
      if (lane == 0 && nx != ~0ull) atomicMin(&next_key, nx);
      __syncthreads();
      key_hi = next_key;
    }
    if (tid == 0) {
      const double xlo = key_to_double(key_lo), xhi = key_to_double(key_hi);
      const double h = tg.h[k];
      double qv = xlo;
      if (h == 1.0) qv = xhi;
      else if (h > 0.0 && h < 1.0 && xlo != xhi)           // no FMA contraction: R rounds both products
        qv = __dadd_rn(__dmul_rn(1.0 - h, xlo), __dmul_rn(h, xhi));
      qtls[(size_t)row * np + k] = qv;
    }
  }
  if (tid == 0) fail[row] = failed ? 1 : 0;
}

static BandTargets make_targets(int64_t N, const double* probs_host, int nprobs) {
  BandTargets tg;
  tg.n = 2 * nprobs;
  const double fuzz = 4 * 2.220446049250313e-16;           // R: fuzz <- 4 * .Machine$double.eps
  for (int k = 0; k < nprobs; ++k) {
    const double nppm = (double)(N - 1) * probs_host[k];   // type 7: (n-1) p
    long long lo = (long long)floor(nppm + fuzz);
    double h = nppm - (double)lo;
    if (fabs(h) < fuzz) h = 0.0;
    if (lo < 0) lo = 0;
    if (lo > N - 1) lo = N - 1;
    long long hi = lo + 1 > N - 1 ? N - 1 : lo + 1;
    tg.rank[2 * k] = lo; tg.rank[2 * k + 1] = hi; tg.h[k] = h;
  }
  return tg;
}

// exact path: transpose into scratch (n_t x N doubles), then one CTA per test point
cudaError_t launch_bands_exact(int n_t, int64_t N, const double* ysim, int64_t ld, const double* probs_host,
                               int nprobs, double* mean, double* qtls, double* scratch, cudaStream_t st) {
  if (nprobs < 0 || 2 * nprobs > MAXT || N <= 0) return cudaErrorInvalidValue;
  const BandTargets tg = make_targets(N, probs_host, nprobs);
  dim3 tb(32, 8), tgid((unsigned)((N + 31) / 32), (unsigned)((n_t + 31) / 32));
  transpose_kernel<<<tgid, tb, 0, st>>>(n_t, N, ysim, ld, scratch, N);
  int grid = n_t < 148 * 4 ? n_t : 148 * 4;
  bands_kernel<<<grid, BT, 0, st>>>(n_t, N, scratch, N, tg, mean, qtls);
  launch_counter() += 2;
  return cudaGetLastError();
}

// resident collect CTAs on the current device for this number of probabilities
static int collect_slots(int nprobs) {
  int per_sm = 0, dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaError_t e = cudaErrorInvalidValue;
#define LQG_OCC(NP_) case NP_: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bands_collect_kernel<NP_>, CROWS, 0); break;
  switch (nprobs) {
    LQG_OCC(0) LQG_OCC(1) LQG_OCC(2) LQG_OCC(3) LQG_OCC(4) LQG_OCC(5) LQG_OCC(6) LQG_OCC(7) LQG_OCC(8)
    default: break;
  }
#undef LQG_OCC
  if (e != cudaSuccess || per_sm < 1) { cudaGetLastError(); per_sm = 1; }
  return sms * per_sm;
}

static BandPlan make_plan(int n_t, int64_t N, const double* probs_host, int nprobs) {
  BandPlan pl;
  memset(&pl, 0, sizeof(pl));
  pl.nprobs = nprobs;
  pl.sampled = N >= 4 * SAMP;
  // enough CTAs to fill the device a few times over: (n_t / 256 tiles) x splits
  const int tiles = (n_t + CROWS - 1) / CROWS;
  int splits = collect_slots(nprobs) / tiles;   // one full wave of resident CTAs (a partial second wave costs a full one)
  if (splits > MAXSPLIT) splits = MAXSPLIT;
  if ((int64_t)splits > N) splits = (int)N;
  if (splits < 1) splits = 1;
  pl.splits = splits;
  const long long per = (N + splits - 1) / splits;          // samples per split
  // bracket width ~ (2 m + 3) / SAMP <= 10 % of the samples; 15 % + slack per split
  pl.cap_split = pl.sampled ? (long long)(0.15 * 1.25 * (double)per) + 256 : per;
  for (int k = 0; k < nprobs; ++k) {
    // sample index e holds element floor(e N / SAMP): the order statistic of rank r sits near sample
    // position r SAMP / N; 4 binomial standard deviations of margin on either side
    const double p = probs_host[k];
    const double q = p * (SAMP - 1);
    const int m = (int)ceil(4.0 * sqrt(SAMP * p * (1.0 - p))) + 4;
    pl.lo_i[k] = (int)floor(q) - m;
    pl.hi_i[k] = (int)ceil(q) + 1 + m;
  }
  return pl;
}

// Scratch (bytes) the bracketed path needs; 0 when it does not apply (N too large for 32-bit counts)
size_t bands_bracket_scratch_bytes(int n_t, int64_t N, int nprobs) {
  if (N >= ((int64_t)1 << 31) || nprobs > MAXP || nprobs < 0) return 0;
  std::vector<double> dummy(nprobs > 0 ? nprobs : 1, 0.5);
  const BandPlan pl = make_plan(n_t, N, dummy.data(), nprobs);
  const int npz = nprobs > 0 ? nprobs : 1;
  size_t b = 0;
  b += (size_t)n_t * npz * 2 * sizeof(uint64_t);                                        // brackets
  b += (size_t)n_t * pl.splits * 2 * sizeof(double);                                    // partial sums + compensation
  b += 2 * (size_t)pl.splits * n_t * npz * sizeof(unsigned);                            // below, counts
  b += (size_t)n_t * sizeof(int);                                                       // fail flags
  b += (size_t)n_t * nprobs * pl.splits * (size_t)pl.cap_split * sizeof(double) + 512;  // buffers
  return b;
}

// bracketed path; fail_dev (n_t ints inside scratch) is returned so that the caller can route the
// flagged test points through launch_bands_exact
cudaError_t launch_bands_bracketed(int n_t, int64_t N, const double* ysim, int64_t ld, const double* probs_host,
                                   int nprobs, double* mean, double* qtls, void* scratch, int** fail_dev,
                                   cudaStream_t st) {
  if (nprobs < 0 || nprobs > MAXP || N <= 0 || N >= ((int64_t)1 << 31)) return cudaErrorInvalidValue;
  const BandTargets tg = make_targets(N, probs_host, nprobs);
  const BandPlan pl = make_plan(n_t, N, probs_host, nprobs);
  const int npz = nprobs > 0 ? nprobs : 1;
  unsigned char* w = static_cast<unsigned char*>(scratch);
  uint64_t* brackets = reinterpret_cast<uint64_t*>(w); w += (size_t)n_t * npz * 2 * sizeof(uint64_t);
  double* part_sum = reinterpret_cast<double*>(w);     w += (size_t)n_t * pl.splits * 2 * sizeof(double);
  unsigned* below_g = reinterpret_cast<unsigned*>(w);  w += (size_t)pl.splits * n_t * npz * sizeof(unsigned);
  unsigned* cnt_g = reinterpret_cast<unsigned*>(w);    w += (size_t)pl.splits * n_t * npz * sizeof(unsigned);
  int* fail = reinterpret_cast<int*>(w);               w += (size_t)n_t * sizeof(int);
  w = reinterpret_cast<unsigned char*>(((uintptr_t)w + 255) & ~(uintptr_t)255);
  double* buf = reinterpret_cast<double*>(w);
  if (pl.sampled && nprobs > 0) {
    bands_sample_kernel<<<n_t, SAMP / 2, 0, st>>>(n_t, N, ysim, ld, pl, brackets);
    ++launch_counter();
  }
  dim3 grid((n_t + CROWS - 1) / CROWS, pl.splits);
#define LQG_COLLECT(NP_) \
  case NP_: bands_collect_kernel<NP_><<<grid, CROWS, 0, st>>>(n_t, N, ysim, ld, pl, brackets, part_sum, below_g, cnt_g, buf); break;
  switch (nprobs) {
    LQG_COLLECT(0) LQG_COLLECT(1) LQG_COLLECT(2) LQG_COLLECT(3) LQG_COLLECT(4)
    LQG_COLLECT(5) LQG_COLLECT(6) LQG_COLLECT(7) LQG_COLLECT(8)
    default: return cudaErrorInvalidValue;
  }
#undef LQG_COLLECT
  bands_select_kernel<<<n_t, BT, 0, st>>>(n_t, N, pl, tg, part_sum, below_g, cnt_g, buf, mean, qtls, fail);
  launch_counter() += 2;
  *fail_dev = fail;
  return cudaGetLastError();
}

// ---- fused form: the product's epilogue collects (dgemm_tma_kernel<true>, lqg_dgemm.cu) ----------------------
// Xs[:, e] = X[:, floor(e N / SAMP)]: the sample columns the brackets are drawn from
__global__ void __launch_bounds__(256)
gather_sample_cols_kernel(int m, int64_t N, const double* __restrict__ X, int64_t ldx, double* __restrict__ Xs, int64_t ldxs) {
  const int e = blockIdx.x;
  const int64_t n = ((int64_t)e * N) / SAMP;
  for (int i = threadIdx.x; i < ldxs; i += blockDim.x) Xs[(size_t)e * ldxs + i] = i < m ? X[(size_t)n * ldx + i] : 0.0;
}

// part_sum[row] = compensated sum of the partial row sums the product left, in a fixed order: a CTA takes 32 test
// points (lane = test point: 256-byte loads), warp w the parts w, w + 8, ...; warp 0 then adds the eight slices in order
__global__ void __launch_bounds__(256)
bands_tile_sum_kernel(int n_t, int parts, const double* __restrict__ tile_sum, double* __restrict__ part_sum) {
  __shared__ double s_sum[8][32], s_comp[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row = blockIdx.x * 32 + lane;
  double sum = 0.0, comp = 0.0;
  auto add = [&](double x) {
    const double t = __dadd_rn(sum, x);
    comp = __dadd_rn(comp, fabs(sum) >= fabs(x) ? __dadd_rn(__dsub_rn(sum, t), x) : __dadd_rn(__dsub_rn(x, t), sum));
    sum = t;
  };
  if (row < n_t) {
    const double* src = tile_sum + row;
    int p = warp;
    for (; p + 8 * (UNR - 1) < parts; p += 8 * UNR) {
      double v[UNR];
#pragma unroll
      for (int q = 0; q < UNR; ++q) v[q] = __ldg(src + (size_t)(p + 8 * q) * n_t);
#pragma unroll
      for (int q = 0; q < UNR; ++q) add(v[q]);
    }
    for (; p < parts; p += 8) add(__ldg(src + (size_t)p * n_t));
  }
  s_sum[warp][lane] = sum; s_comp[warp][lane] = comp;
  __syncthreads();
  if (warp == 0 && row < n_t) {
    sum = 0.0; comp = 0.0;
    for (int w = 0; w < 8; ++w) { add(s_sum[w][lane]); add(s_comp[w][lane]); }
    part_sum[(size_t)row * 2] = sum;
    part_sum[(size_t)row * 2 + 1] = comp;
  }
}

namespace {
struct FusedLayout {
  size_t xs, ys, brackets, part_sum, counts, fail, tile_sum, buf, total;
  int64_t ldxs;
  int parts;
  long long cap;
};
FusedLayout fused_layout(int n_t, int m, int64_t N, int nprobs) {
  FusedLayout L;
  const int npz = nprobs > 0 ? nprobs : 1;
  L.ldxs = m + (m & 1);                                   // even leading dimension: 16-byte aligned columns for the tensor map
  L.parts = dgemm_bands_parts((int)N);
  L.cap = (long long)(0.15 * 1.25 * (double)N) + 256;     // as make_plan: brackets hold <= 10 % of a row
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t at = off; off = (off + bytes + 255) & ~(size_t)255; return at; };
  L.xs = take((size_t)L.ldxs * SAMP * sizeof(double));
  L.ys = take((size_t)n_t * SAMP * sizeof(double));
  L.brackets = take((size_t)n_t * npz * 2 * sizeof(uint64_t));
  L.part_sum = take((size_t)n_t * 2 * sizeof(double));
  L.counts = take(2 * (size_t)n_t * npz * sizeof(unsigned));        // below | cursor
  L.fail = take((size_t)n_t * sizeof(int));
  L.tile_sum = take((size_t)L.parts * n_t * sizeof(double));
  L.buf = take((size_t)n_t * nprobs * (size_t)L.cap * sizeof(double) + 512);
  L.total = off;
  return L;
}
}  // namespace

size_t paths_bands_fused_scratch_bytes(int n_t, int m, int64_t N, int nprobs, const double* phi, int64_t ldphi,
                                       const double* X, int64_t ldx) {
  if (N < 4 * SAMP || N >= ((int64_t)1 << 31) || nprobs > MAXP || nprobs < 0 || n_t <= 0 || m <= 0) return 0;
  if (!dgemm_bands_usable(phi, ldphi, X, ldx)) return 0;
  return fused_layout(n_t, m, N, nprobs).total;
}

cudaError_t launch_paths_bands_fused(int n_t, int m, int64_t N, const double* phi, int64_t ldphi, const double* X, int64_t ldx,
                                     const double* probs_host, int nprobs, double* mean, double* qtls, void* scratch,
                                     int** fail_dev, cudaStream_t st, cudaEvent_t ev_product_done) {
  if (paths_bands_fused_scratch_bytes(n_t, m, N, nprobs, phi, ldphi, X, ldx) == 0) return cudaErrorNotSupported;
  const FusedLayout L = fused_layout(n_t, m, N, nprobs);
  const BandTargets tg = make_targets(N, probs_host, nprobs);
  BandPlan pl = make_plan(n_t, N, probs_host, nprobs);      // bracket positions; one segment per (test point, probability)
  pl.splits = 1;
  pl.cap_split = L.cap;
  const int npz = nprobs > 0 ? nprobs : 1;
  unsigned char* w = static_cast<unsigned char*>(scratch);
  double* xs = reinterpret_cast<double*>(w + L.xs);
  double* ys = reinterpret_cast<double*>(w + L.ys);
  uint64_t* brackets = reinterpret_cast<uint64_t*>(w + L.brackets);
  double* part_sum = reinterpret_cast<double*>(w + L.part_sum);
  unsigned* below = reinterpret_cast<unsigned*>(w + L.counts);
  unsigned* cursor = below + (size_t)n_t * npz;
  int* fail = reinterpret_cast<int*>(w + L.fail);
  cudaError_t e;
  if (nprobs > 0) {
    // pilot: the 2048 sampled columns of the product, sorted per test point -> brackets
    gather_sample_cols_kernel<<<SAMP, 256, 0, st>>>(m, N, X, ldx, xs, L.ldxs);
    ++launch_counter();
    if ((e = launch_dgemm(n_t, SAMP, m, phi, ldphi, xs, L.ldxs, ys, n_t, nullptr, 0, st)) != cudaSuccess) return e;
    bands_sample_kernel<<<n_t, SAMP / 2, 0, st>>>(n_t, SAMP, ys, n_t, pl, brackets);
    ++launch_counter();
  }
  if ((e = cudaMemsetAsync(below, 0, 2 * (size_t)n_t * npz * sizeof(unsigned), st)) != cudaSuccess) return e;
  BandFuse f;
  f.brackets = reinterpret_cast<const unsigned long long*>(brackets);
  f.below = below; f.cursor = cursor;
  f.buf = reinterpret_cast<double*>(w + L.buf);
  f.tile_sum = reinterpret_cast<double*>(w + L.tile_sum);
  f.cap = (unsigned)L.cap;
  f.nprobs = nprobs;
  if ((e = launch_dgemm_bands(n_t, (int)N, m, phi, ldphi, X, ldx, f, st)) != cudaSuccess) return e;
  if (ev_product_done && (e = cudaEventRecord(ev_product_done, st)) != cudaSuccess) return e;
  bands_tile_sum_kernel<<<(n_t + 31) / 32, 256, 0, st>>>(n_t, L.parts, f.tile_sum, part_sum);
  bands_select_kernel<<<n_t, BT, 0, st>>>(n_t, N, pl, tg, part_sum, below, cursor, f.buf, mean, qtls, fail);
  launch_counter() += 2;
  *fail_dev = fail;
  return cudaGetLastError();
}

}  // namespace lqg
