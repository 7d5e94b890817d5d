// FP64 roofline denominators measured on the device itself (MEASURED_PEAKS.json has HBM and
// bf16 only): a DFMA-issue micro-benchmark for the CUDA-core FP64 pipe (bounds the HMC chain
// kernels) and a DMMA.8x8x4 micro-benchmark for the FP64 tensor path (bounds lqg_dgemm).
#include "lqg_common.cuh"
#include "lqg_kernels.h"

namespace lqg {

__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double x) {
  double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double y = x * 0.5;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, x, y); a1 = fma(a1, x, y); a2 = fma(a2, x, y); a3 = fma(a3, x, y);
    a4 = fma(a4, x, y); a5 = fma(a5, x, y); a6 = fma(a6, x, y); a7 = fma(a7, x, y);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters, double x) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x + i; c[i][1] = -c[i][0]; }
  const double a = x, b = x * 0.25;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Do the FP64 FMA pipe and the FP64 tensor path add up?  Even warps issue DFMA, odd warps DMMA, both
// with the per-warp work of the kernels above.  If the sum of the two rates exceeds either peak the
// units are separate; on B200 it does not (see profiles/): DMMA runs on the same 64 FMA/clk/SM.
__global__ void __launch_bounds__(256) fp64_mixed_kernel(double* out, int iters_f, int iters_m, double x) {
  const int warp = threadIdx.x >> 5;
  double s = 0;
  if (warp & 1) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x + i; c[i][1] = -c[i][0]; }
    const double a = x, b = x * 0.25;
    for (int it = 0; it < iters_m; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  } else {
    double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double y = x * 0.5;
    for (int i = 0; i < iters_f; ++i) {
      a0 = fma(a0, x, y); a1 = fma(a1, x, y); a2 = fma(a2, x, y); a3 = fma(a3, x, y);
      a4 = fma(a4, x, y); a5 = fma(a5, x, y); a6 = fma(a6, x, y); a7 = fma(a7, x, y);
    }
    s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// combined Tflop/s of the mixed kernel (DFMA warps + DMMA warps running side by side)
cudaError_t measure_fp64_mixed(double* tflops, double* dfma_share, cudaStream_t st) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = sms * 8, nt = 256;
  double* buf = nullptr;
  cudaError_t e = cudaMalloc(&buf, (size_t)grid * nt * sizeof(double));
  if (e != cudaSuccess) return e;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  // per-warp iteration counts sized so that each half would take about the same time alone
  const int it_f = 1 << 15, it_m = 1 << 12;      // DFMA: 16 flop x 32 lanes per iteration; DMMA: 8 x 512 flop per iteration
  float ms = 0;
  double best = 0;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(a, st);
    fp64_mixed_kernel<<<grid, nt, 0, st>>>(buf, it_f, it_m, 0.999999);
    cudaEventRecord(b, st);
    cudaEventSynchronize(b);
    cudaEventElapsedTime(&ms, a, b);
    ++launch_counter();
    const double ff = 2.0 * 8.0 * it_f * (double)grid * (nt / 2);
    const double fm = 512.0 * 8.0 * it_m * (double)grid * (nt / 64);
    const double tf = (ff + fm) / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) { best = tf; if (dfma_share) *dfma_share = ff / (ff + fm); }
  }
  cudaEventDestroy(a); cudaEventDestroy(b);
  e = cudaGetLastError();
  cudaFree(buf);
  if (tflops) *tflops = best;
  return e;
}

cudaError_t measure_fp64_peaks(double* dfma_tflops, double* dmma_tflops, cudaStream_t st) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = sms * 8, nt = 256;
  double* buf = nullptr;
  cudaError_t e = cudaMalloc(&buf, (size_t)grid * nt * sizeof(double));
  if (e != cudaSuccess) return e;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  float ms = 0;
  double best_f = 0, best_m = 0;
  const int it_f = 1 << 15, it_m = 1 << 12;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(a, st);
    dfma_peak_kernel<<<grid, nt, 0, st>>>(buf, it_f, 0.999999);
    cudaEventRecord(b, st);
    cudaEventSynchronize(b);
    cudaEventElapsedTime(&ms, a, b);
    ++launch_counter();
    double tf = 2.0 * 8.0 * it_f * (double)grid * nt / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best_f) best_f = tf;
    cudaEventRecord(a, st);
    dmma_peak_kernel<<<grid, nt, 0, st>>>(buf, it_m, 0.999999);
    cudaEventRecord(b, st);
    cudaEventSynchronize(b);
    cudaEventElapsedTime(&ms, a, b);
    ++launch_counter();
    // one m8n8k4 = 8*8*4*2 flop per warp
    double tm = 512.0 * 8.0 * it_m * (double)grid * (nt / 32) / (ms * 1e-3) / 1e12;
    if (rep > 0 && tm > best_m) best_m = tm;
  }
  cudaEventDestroy(a); cudaEventDestroy(b);
  e = cudaGetLastError();
  cudaFree(buf);
  if (dfma_tflops) *dfma_tflops = best_f;
  if (dmma_tflops) *dmma_tflops = best_m;
  return e;
}

}  // namespace lqg
