// Ordered-compaction helpers shared by the rejection samplers (RSM, ExpT): a 256-thread block
// exclusive scan (shuffle inside the warp, one slot per warp) and a single-CTA scan of the
// per-block counts.
#pragma once
#include "lqg_common.cuh"

namespace lqg {

constexpr int kScanBlock = 256;

__device__ __forceinline__ int block_excl_scan(int v, int* warp_tot, int* total) {
  // exclusive scan over the 256 threads of a block
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    int o = __shfl_up_sync(0xffffffffu, inc, off);
    if (lane >= off) inc += o;
  }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kScanBlock / 32; ++w) { const int t = warp_tot[w]; if (w < warp) base += t; tot += t; }
  __syncthreads();
  *total = tot;
  return base + inc - v;
}


// exclusive scan of n ints by one CTA of 1024 threads; off[n] = total
static __global__ void __launch_bounds__(1024) scan_counts_kernel(const int* in, int64_t* off, int n) {
  __shared__ int64_t wt[32];
  __shared__ int64_t carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    int64_t v = i < n ? in[i] : 0, inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int64_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) wt[warp] = inc;
    __syncthreads();
    int64_t wb = 0, tot = 0;
    for (int w = 0; w < 32; ++w) { if (w < warp) wb += wt[w]; tot += wt[w]; }
    const int64_t c = carry;
    if (i < n) off[i] = c + wb + inc - v;
    __syncthreads();
    if (threadIdx.x == 0) carry = c + tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) off[n] = carry;
}


}  // namespace lqg
