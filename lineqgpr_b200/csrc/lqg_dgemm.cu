// FP64 tensor-core GEMM  C (m x n) = A (m x k) * B (k x n), everything column-major.
//
// Users (all dense FP64 contractions on the hot path):
//   * momentum pre-pass  X_a = U A   (d x d  times  d x #transitions):  the whitening product
//     that tmg folds into f2 = f %*% Ri (tmg:R/tmg.R:50), batched over every chain/transition
//   * xi.sim = Lambda^+ eta          R/lineqGP.R:639, R/lineqAGP.R:575
//   * ysim   = Phi.test %*% xi.sim   R/lineqGP.R:640, R/lineqAGP.R:509-512
//
// tcgen05/UMMA has no FP64 kind, so the FP64 tensor path on sm_100a is
// mma.sync.aligned.m8n8k4.f64 (SASS: DMMA.8x8x4).  CTA tile 128 x 64 x 16, 8 warps as 4 x 2,
// warp tile 32 x 32 = 4 x 4 DMMA fragments; operands staged through padded shared memory
// (conflict-free fragment reads), register-prefetch double buffering, one barrier per k-tile.
// For an upper-triangular A (tmg's Ri) the all-zero k-tiles below the diagonal are skipped.
#include <stdlib.h>

#include <algorithm>

#include "lqg_common.cuh"
#include "lqg_kernels.h"

namespace lqg {

namespace {

constexpr int BM = 128, BN = 64, BK = 16;
constexpr int NT = 256;
constexpr int A_LD = BM + 4;     // (k*A_LD + row): per half-warp 4 k x 4 rows hit 16 distinct 8-byte banks
constexpr int B_LD = BK + 4;     // (n*B_LD + k):   per half-warp 4 n x 4 k hit 16 distinct 8-byte banks

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

}  // namespace

__global__ void __launch_bounds__(NT)
dgemm_dmma_kernel(int m, int n, int k, const double* __restrict__ A, int64_t lda,
                  const double* __restrict__ B, int64_t ldb, double* __restrict__ C, int64_t ldc,
                  double* __restrict__ row_sum, int a_upper, int b_blk_w, int64_t b_blk_pitch, int tail_cb) {
  extern __shared__ __align__(16) double dg_smem[];
  double (*As)[BK * A_LD] = reinterpret_cast<double (*)[BK * A_LD]>(dg_smem);
  double (*Bs)[BN * B_LD] = reinterpret_cast<double (*)[BN * B_LD]>(dg_smem + 2 * BK * A_LD);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 3, wn = warp >> 2;
  // Tile order (1-D grid): row block fastest, so that the CTAs running side by side share one B tile
  // and differ in length (their prologues and epilogues do not line up).  With an upper-triangular A a
  // top row block has 25 x the k-tiles of a bottom one: if the last column blocks were issued in that
  // order too, a few long tiles would be left running alone at the end of the launch (~9 % of it).
  // The last `tail_cb` column blocks are therefore issued row block by row block, longest first.
  const int MB = (m + BM - 1) / BM, NB = (n + BN - 1) / BN;
  const int NBh = NB - tail_cb;
  int bm, bn;
  if ((int)blockIdx.x < MB * NBh) { bn = blockIdx.x / MB; bm = blockIdx.x - bn * MB; }
  else { const int r = blockIdx.x - MB * NBh; bm = r / tail_cb; bn = NBh + (r - bm * tail_cb); }
  const int row0 = bm * BM, col0 = bn * BN;

  // A tile loader: thread -> (row = tid % 128, k in {kh*8 .. kh*8+7})
  const int a_row = tid & (BM - 1), a_kh = tid >> 7;
  // B tile loader: thread -> (k = tid % 16, n = tid/16 + 16 q)
  const int b_k = tid & (BK - 1), b_n = tid >> 4;

  double ra[8], rb[4];
  // B columns may come in blocks (b_blk_w > 0): column j lives at B + (j / w) * pitch + (j % w) * ldb.
  // That is how a round of chain-major samples looks: w finished columns of every chain.
  const double* bcol[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int gc = col0 + b_n + 16 * q;
    bcol[q] = gc >= n ? nullptr
            : b_blk_w > 0 ? B + (size_t)(gc / b_blk_w) * b_blk_pitch + (size_t)(gc % b_blk_w) * ldb
                          : B + (size_t)gc * ldb;
  }
  const int kt_end = (k + BK - 1) / BK;
  // upper-triangular A: A[i, kk] = 0 for i > kk, so k-tiles entirely left of the row block vanish
  const int kt_begin = a_upper ? min(row0 / BK, kt_end) : 0;

  auto load_tile = [&](int kt) {
    const int k0 = kt * BK;
    const int gr = row0 + a_row;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int kk = k0 + a_kh * 8 + q;
      ra[q] = (gr < m && kk < k) ? __ldg(A + (size_t)kk * lda + gr) : 0.0;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int kk = k0 + b_k;
      rb[q] = (bcol[q] && kk < k) ? __ldg(bcol[q] + kk) : 0.0;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int q = 0; q < 8; ++q) As[buf][(a_kh * 8 + q) * A_LD + a_row] = ra[q];
#pragma unroll
    for (int q = 0; q < 4; ++q) Bs[buf][(b_n + 16 * q) * B_LD + b_k] = rb[q];
  };

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  if (kt_begin < kt_end) {
    load_tile(kt_begin);
    store_tile(0);
  }
  __syncthreads();

  const int fr = lane >> 2, fk = lane & 3;
  int buf = 0;
  for (int kt = kt_begin; kt < kt_end; ++kt) {
    if (kt + 1 < kt_end) load_tile(kt + 1);
    const double* as = As[buf];
    const double* bs = Bs[buf];
#pragma unroll
    for (int k4 = 0; k4 < BK / 4; ++k4) {
      double af[4], bf[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) af[i] = as[(k4 * 4 + fk) * A_LD + wm * 32 + i * 8 + fr];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = bs[(wn * 32 + j * 8 + fr) * B_LD + k4 * 4 + fk];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
    if (kt + 1 < kt_end) store_tile(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }

  // ---- epilogue: C store and optional row sums (band mean numerator) ----
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = row0 + wm * 32 + i * 8 + fr;
    double rs = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int cc = col0 + wn * 32 + j * 8 + 2 * fk;
      if (r < m) {
        if (cc < n)     { if (C) C[(size_t)cc * ldc + r] = acc[i][j][0];       rs += acc[i][j][0]; }
        if (cc + 1 < n) { if (C) C[(size_t)(cc + 1) * ldc + r] = acc[i][j][1]; rs += acc[i][j][1]; }
      }
    }
    if (row_sum) {
      rs += __shfl_xor_sync(0xffffffffu, rs, 1);
      rs += __shfl_xor_sync(0xffffffffu, rs, 2);
      if (fk == 0 && r < m) atomicAdd(row_sum + r, rs);
    }
  }
}

// A[kk, col] = N(0,1) draw of (chain = active[col / attempts], draw index = pos[chain] + col % attempts,
// coordinate kk) — the same Philox addressing the chain kernels use, so pool and in-kernel
// draws agree.
__global__ void fill_normals_kernel(int d, size_t n_cols, double* __restrict__ A, NormalGen g) {
  const int pairs = (d + 1) / 2;
  const size_t total = (size_t)pairs * n_cols;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const uint32_t pr = (uint32_t)(idx % pairs);
    const size_t col = idx / pairs;
    const int a = (int)(col / g.attempts);
    const int chain = g.active ? g.active[a] : a;
    const uint64_t draw = (uint64_t)(g.pos[chain] + (int64_t)(col % g.attempts));
    double n0, n1;
    normal_pair(g.seed, (uint64_t)(g.chain_offset + chain), draw, pr, n0, n1);
    double* dst = A + col * d + 2 * (size_t)pr;
    dst[0] = n0;
    if (2 * pr + 1 < (uint32_t)d) dst[1] = n1;
  }
}

cudaError_t launch_dgemm(int m, int n, int k, const double* A, int64_t lda, const double* B,
                         int64_t ldb, double* C, int64_t ldc, double* row_sum, int a_upper,
                         cudaStream_t st, int b_blk_w, int64_t b_blk_pitch) {
  if (m <= 0 || n <= 0) return cudaSuccess;
  const size_t smem = (size_t)(2 * BK * A_LD + 2 * BN * B_LD) * sizeof(double);
  cudaError_t e0 = cudaFuncSetAttribute(dgemm_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e0 != cudaSuccess) return e0;
  const long long tiles = (long long)((m + BM - 1) / BM) * ((n + BN - 1) / BN);
  if (tiles > 0x7fffffffLL) return cudaErrorInvalidValue;
  int tail_cb = 0;
  if (a_upper) {                       // four waves of resident CTAs' worth of column blocks, longest tiles first
    static const int slots = [] {
      int dev = 0, sms = 148, per_sm = 2;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dgemm_dmma_kernel, NT, (size_t)(2 * BK * A_LD + 2 * BN * B_LD) * sizeof(double));
      return sms * (per_sm > 0 ? per_sm : 1);
    }();
    static const char* env_tail = getenv("LQG_GEMM_TAIL");           // tuning aid: waves of column blocks reordered (default 4)
    const double waves = env_tail ? atof(env_tail) : 4.0;
    const int MBc = (m + BM - 1) / BM, NBc = (n + BN - 1) / BN;
    tail_cb = std::min(NBc, (int)(waves * slots / MBc));
  }
  dgemm_dmma_kernel<<<(unsigned)tiles, NT, smem, st>>>(m, n, k, A, lda, B, ldb, C, ldc, row_sum, a_upper, b_blk_w, b_blk_pitch, tail_cb);
  ++launch_counter();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  return cudaSuccess;
}

cudaError_t launch_fill_normals(int d, size_t n_cols, double* A, const NormalGen& g, cudaStream_t st) {
  fill_normals_kernel<<<148 * 8, 256, 0, st>>>(d, n_cols, A, g);
  ++launch_counter();
  return cudaGetLastError();
}

}  // namespace lqg
