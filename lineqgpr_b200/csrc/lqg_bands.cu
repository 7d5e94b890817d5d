// Summary bands of the simulated paths: row means and R type-7 quantiles of ysim (n_t x N),
//   qtls  <- apply(ysim, 1, quantile, probs = probs)      R/lineqGPutils.R:422, :296
//   ymean <- apply(ysim, 1, mean)                          R/lineqGPutils.R:423, :294
//
// ysim is column-major with one COLUMN per sample, so the N samples of one test point are
// strided; a tiled transpose first makes every test point's samples contiguous.  One CTA then
// owns one test point: a deterministic tree sum for the mean, and an EXACT order-statistic
// search for the 2*nprobs ranks a type-7 quantile needs — most-significant-digit radix select
// (8 passes of 8 bits over an order-preserving 64-bit key, per-target histograms in shared
// memory, warp-aggregated atomics), no sorting and no approximation.
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

cudaError_t launch_bands(int n_t, int64_t N, const double* ysim, int64_t ld, const double* probs_host,
                         int nprobs, double* mean, double* qtls, double* scratch, cudaStream_t st) {
  if (nprobs < 0 || 2 * nprobs > MAXT || N <= 0) return cudaErrorInvalidValue;
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
  dim3 tb(32, 8), tgid((unsigned)((N + 31) / 32), (unsigned)((n_t + 31) / 32));
  transpose_kernel<<<tgid, tb, 0, st>>>(n_t, N, ysim, ld, scratch, N);
  int grid = n_t < 148 * 4 ? n_t : 148 * 4;
  bands_kernel<<<grid, BT, 0, st>>>(n_t, N, scratch, N, tg, mean, qtls);
  launch_counter() += 2;
  return cudaGetLastError();
}

}  // namespace lqg
