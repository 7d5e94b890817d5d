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
//
// Two kernels share the tile shape and the epilogue:
//   * dgemm_ring_kernel (default): operands move global -> shared memory with bulk-async copies
//     (cp.async.bulk, the TMA engine; SASS UBLKCP) into a 4-stage ring guarded by full/empty
//     mbarriers.  There is no CTA-wide barrier in the k loop and no operand passes through
//     registers: every warp issues its share of the copies two k-tiles ahead (2 columns of A,
//     8 columns of B) and otherwise only reads fragments and issues DMMA.  Inside the diagonal
//     block of an upper-triangular A the 8-row fragment groups that lie entirely below the
//     diagonal are skipped as well.  Needs 16-byte aligned operand columns (even leading
//     dimensions); partial k-tiles and odd row counts are staged by the threads themselves.
//   * dgemm_dmma_kernel: LDG -> registers -> STS double buffering with one barrier per k-tile;
//     kept for operands that are not 16-byte aligned (LQG_GEMM_RING=0 forces it).
#include <cuda.h>
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


// ---- ring kernel -------------------------------------------------------------------------
constexpr int ST = 4;                         // ring stages
constexpr int PF = 2;                         // k-tiles in flight ahead of the one being multiplied
constexpr int A_STAGE = BK * A_LD;            // doubles per stage: A tile, column kk at kk * A_LD
constexpr int B_STAGE = BN * B_LD;            //                    B tile, column n at n * B_LD
constexpr int STAGE = A_STAGE + B_STAGE;
constexpr int NWARP = NT / 32;
static_assert(BK == 2 * NWARP && BN == 8 * NWARP, "every warp stages 2 columns of A and 8 columns of B");
static_assert((A_LD * 8) % 16 == 0 && (B_LD * 8) % 16 == 0 && (A_STAGE * 8) % 16 == 0 && (STAGE * 8) % 16 == 0,
              "bulk copies need 16-byte aligned shared-memory destinations");

__global__ void __launch_bounds__(NT, 2)
dgemm_ring_kernel(int m, int n, int k, const double* __restrict__ A, int64_t lda,
                  const double* __restrict__ B, int64_t ldb, double* __restrict__ C, int64_t ldc,
                  double* __restrict__ row_sum, int a_upper, int b_blk_w, int64_t b_blk_pitch, int tail_cb) {
  extern __shared__ __align__(128) double dg_ring[];
  __shared__ __align__(8) uint64_t full_bar[ST], empty_bar[ST];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 3, wn = warp >> 2;
  const int MB = (m + BM - 1) / BM, NB = (n + BN - 1) / BN;     // tile order: see dgemm_dmma_kernel
  const int NBh = NB - tail_cb;
  int bm, bn;
  if ((int)blockIdx.x < MB * NBh) { bn = blockIdx.x / MB; bm = blockIdx.x - bn * MB; }
  else { const int r = blockIdx.x - MB * NBh; bm = r / tail_cb; bn = NBh + (r - bm * tail_cb); }
  const int row0 = bm * BM, col0 = bn * BN;
  const int rows_valid = min(BM, m - row0);
  const int kt_end = (k + BK - 1) / BK;
  const int kt_begin = a_upper ? min(row0 / BK, kt_end) : 0;
  const int nkt = kt_end - kt_begin;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < ST; ++s) { mbar_init(&full_bar[s], NWARP); mbar_init(&empty_bar[s], NWARP); }
  }
  __syncthreads();

  auto b_col = [&](int gc) -> const double* {                   // global column gc of B (see b_blk_w)
    return b_blk_w > 0 ? B + (size_t)(gc / b_blk_w) * b_blk_pitch + (size_t)(gc % b_blk_w) * ldb
                       : B + (size_t)gc * ldb;
  };
  // the copies this lane issues for a full k-tile: lanes 0,1 one column of A each, lanes 2..9 one column of B each
  const int my_b = col0 + 8 * warp + (lane - 2);
  const double* my_src = lane < 2 ? A + (size_t)(2 * warp + lane) * lda + row0
                       : (lane < 10 && my_b < n) ? b_col(my_b) : nullptr;
  const int my_dst = lane < 2 ? (2 * warp + lane) * A_LD : A_STAGE + (8 * warp + (lane - 2)) * B_LD;
  const uint32_t my_bytes = lane < 2 ? (uint32_t)rows_valid * 8u : (uint32_t)BK * 8u;
  const int64_t my_kstride = lane < 2 ? lda : 1;                // elements per k step in this lane's source
  const int b_valid = max(0, min(8, n - col0 - 8 * warp));      // valid columns among the warp's 8
  const uint32_t warp_bytes = 2u * (uint32_t)rows_valid * 8u + (uint32_t)b_valid * BK * 8u;
  const bool rows_even = (rows_valid & 1) == 0;

  // stage k-tile `it` (counted from kt_begin) into ring slot it % ST
  auto issue = [&](int it) {
    const int s = it % ST;
    const int k0 = (kt_begin + it) * BK;
    double* stage = dg_ring + (size_t)s * STAGE;
    if (it >= ST) mbar_wait(&empty_bar[s], (uint32_t)((it / ST) - 1) & 1u);   // every warp is done with the slot's previous tile
    if (rows_even && k0 + BK <= k) {
      if (lane == 0) mbar_expect_tx(&full_bar[s], warp_bytes);
      __syncwarp();
      if (my_src) bulk_g2s(stage + my_dst, my_src + (size_t)k0 * my_kstride, my_bytes, &full_bar[s]);
    } else {
      // partial k-tile (or an odd number of rows): plain loads, zero fill
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int kk = k0 + 2 * warp + c;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int r = lane + 32 * q;
          stage[(2 * warp + c) * A_LD + r] = (row0 + r < m && kk < k) ? __ldg(A + (size_t)kk * lda + row0 + r) : 0.0;
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int nn = 8 * warp + (lane >> 4) + 2 * q, kk = k0 + (lane & 15);
        const int gc = col0 + nn;
        stage[A_STAGE + nn * B_LD + (lane & 15)] = (gc < n && kk < k) ? __ldg(b_col(gc) + kk) : 0.0;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[s]);
    }
  };

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  for (int it = 0; it < PF && it < nkt; ++it) issue(it);

  const int fr = lane >> 2, fk = lane & 3;
  const int wrow = row0 + wm * 32;                                // first row of the warp tile
  for (int it = 0; it < nkt; ++it) {
    if (it + PF < nkt) issue(it + PF);
    const int s = it % ST;
    const int k0 = (kt_begin + it) * BK;
    mbar_wait(&full_bar[s], (uint32_t)(it / ST) & 1u);
    const double* as = dg_ring + (size_t)s * STAGE + wm * 32 + fr;
    const double* bs = dg_ring + (size_t)s * STAGE + A_STAGE + (wn * 32 + fr) * B_LD + fk;
    if (a_upper && k0 < wrow + 32) {
      // diagonal block of an upper-triangular A: A[i, kk] = 0 for i > kk, so a fragment group of 8 rows
      // starting at r contributes nothing to the 4 columns kk..kk+3 when r > kk + 3
#pragma unroll
      for (int k4 = 0; k4 < BK / 4; ++k4) {
        const int kmax = k0 + k4 * 4 + 3;
        if (wrow > kmax) continue;
        double bf[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) bf[j] = bs[j * 8 * B_LD + k4 * 4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (wrow + i * 8 > kmax) continue;
          const double af = as[(k4 * 4 + fk) * A_LD + i * 8];
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af, bf[j]);
        }
      }
    } else {
#pragma unroll
      for (int k4 = 0; k4 < BK / 4; ++k4) {
        double af[4], bf[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) af[i] = as[(k4 * 4 + fk) * A_LD + i * 8];
#pragma unroll
        for (int j = 0; j < 4; ++j) bf[j] = bs[j * 8 * B_LD + k4 * 4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
      }
    }
    __syncwarp();                                                 // every lane has read its fragments
    if (lane == 0) mbar_arrive(&empty_bar[s]);
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


// ---- tensor-map kernel -------------------------------------------------------------------
// Same tile shape, ring and epilogue as dgemm_ring_kernel, but a k-tile arrives as 9 tiled TMA copies
// (cp.async.bulk.tensor.2d, SASS UTMALDG) issued by ONE thread: 8 boxes of 16 rows x 16 k of A and one box
// of 16 k x 64 columns of B, 128-byte swizzled.  Out-of-range rows, columns and k are zero-filled by the
// copy engine, so there is no edge path.  Shared-memory layout of a stage (1024-byte aligned):
//   A box b (rows 16 b .. 16 b + 15):  byte 2048 b + 128 k + 8 r, with address bits [4:6] ^= bits [7:9]
//   B:                                 byte 16384 + 128 n + 8 k,  same swizzle
// DMMA sums 4 values of k per instruction; which 4 is free as long as A and B agree.  Step k4 of a k-tile
// takes k = 2 q + 4 (fk / 2) + (fk mod 2), q in {0, 1, 4, 5}.  A 64-bit shared-memory load is served half-warp
// by half-warp; with this choice each half-warp of either fragment load covers 8 of the 16 bank pairs twice
// (4 wavefronts per load, ncu: half of the wavefronts are conflicts), the balanced point: a k order that
// makes the A loads conflict free (k bits 1,2 from fk) makes the B loads 4-way conflicting and vice versa,
// because the swizzle XORs the row (A) / column (B) index into the same address bits.  The padded layout of
// the bulk-copy ring needs 2 wavefronts per load, but TMA cannot write padding; at 47 % of the
// shared-memory data pipe the loads are not what limits the kernel (DMMA sub-pipe 94 % active).
constexpr int TST = 4, TPF = 2;
constexpr int T_A_BYTES = BM * BK * 8, T_B_BYTES = BN * BK * 8, T_STAGE_BYTES = T_A_BYTES + T_B_BYTES;
static_assert(T_STAGE_BYTES % 1024 == 0, "stages must keep the 1024-byte alignment of the swizzle pattern");

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}

// bracket end (order-preserving key, lqg_bands.cu ordered_key) back to the double it came from; the open ends
// (key 0 / ~0) become -inf / +inf.  The epilogue compares doubles, not keys: one DSETP instead of a key
// construction and a 64-bit integer compare per value.  The two orders differ only on -0.0 < +0.0, where the
// double comparison makes the bracket one zero wider at that end — still a contiguous range of the key order with
// everything "below" in front of it, which is all the selection relies on.
__device__ __forceinline__ double band_bound(unsigned long long key) {
  if (key == 0ull) return __longlong_as_double(0xfff0000000000000ll);
  if (key == ~0ull) return __longlong_as_double(0x7ff0000000000000ll);
  const unsigned long long b = (key & 0x8000000000000000ull) ? (key & 0x7fffffffffffffffull) : ~key;
  return __longlong_as_double((long long)b);
}

// FUSE: the tile is summarised for the band statistics instead of being stored (BandFuse, lqg_kernels.h) — per
// test point (row) and column half-tile a partial sum, per (row, probability) the count of values below the
// bracket and the values inside it, appended at a cursor.  Integer atomics only: the counts and the SET of
// collected values do not depend on the order the tiles retire in, so the selected order statistics are
// reproducible bit for bit; the sums are reduced afterwards in tile order.
template <bool FUSE>
__global__ void __launch_bounds__(NT, 2)
dgemm_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int m, int n, int k,
                 double* __restrict__ C, int64_t ldc, double* __restrict__ row_sum, int a_upper, int tail_cb, BandFuse bf) {
  extern __shared__ __align__(1024) unsigned char dg_tma[];
  __shared__ __align__(8) uint64_t full_bar[TST], empty_bar[TST];
  // dynamic shared memory is only guaranteed 16-byte aligned: round up (the launch asks for 1 KB more)
  unsigned char* ring = dg_tma + ((1024u - (smem_u32(dg_tma) & 1023u)) & 1023u);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 3, wn = warp >> 2;
  const int MB = (m + BM - 1) / BM, NB = (n + BN - 1) / BN;     // tile order: see dgemm_dmma_kernel
  const int NBh = NB - tail_cb;
  int bm, bn;
  if ((int)blockIdx.x < MB * NBh) { bn = blockIdx.x / MB; bm = blockIdx.x - bn * MB; }
  else { const int r = blockIdx.x - MB * NBh; bm = r / tail_cb; bn = NBh + (r - bm * tail_cb); }
  const int row0 = bm * BM, col0 = bn * BN;
  const int kt_end = (k + BK - 1) / BK;
  const int kt_begin = a_upper ? min(row0 / BK, kt_end) : 0;
  const int nkt = kt_end - kt_begin;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < TST; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], NWARP); }
  }
  __syncthreads();

  // k-tile `it` is issued by lane 0 of warp it % 8 (the job rotates so that no warp's DMMA stream pays for all of it)
  auto issue = [&](int it) {
    if (warp != (it & (NWARP - 1))) return;
    const int s = it % TST;
    if (it >= TST) mbar_wait(&empty_bar[s], (uint32_t)((it / TST) - 1) & 1u);
    if (lane == 0) {
      const int k0 = (kt_begin + it) * BK;
      unsigned char* stage = ring + (size_t)s * T_STAGE_BYTES;
      mbar_expect_tx(&full_bar[s], (uint32_t)T_STAGE_BYTES);
#pragma unroll
      for (int b = 0; b < BM / 16; ++b) tma_load_2d(stage + b * 2048, &mapA, row0 + 16 * b, k0, &full_bar[s]);
      tma_load_2d(stage + T_A_BYTES, &mapB, k0, col0, &full_bar[s]);
    }
  };

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  for (int it = 0; it < TPF && it < nkt; ++it) issue(it);

  const int fr = lane >> 2, fk = lane & 3;
  // lane parts of the swizzled fragment addresses (bytes; see the layout above)
  const int a_lane = wm * 4096 + (4 * (fk >> 1) + (fk & 1)) * 128 + 8 * (fr & 1);
  const int a_c0 = (fr >> 1) ^ (fk & 1);
  const int a_off[2][2] = {{64 * (fk >> 1) + 16 * a_c0, 64 * (fk >> 1) + 16 * (a_c0 ^ 2)},
                           {64 * (1 - (fk >> 1)) + 16 * a_c0, 64 * (1 - (fk >> 1)) + 16 * (a_c0 ^ 2)}};   // [i & 1][k4 & 1]
  const int b_lane = T_A_BYTES + (wn * 32 + fr) * 128 + 8 * (fk & 1);
  const int b_l = ((fk >> 1) << 1) ^ fr;
  const int b_off[4] = {16 * (0 ^ b_l), 16 * (1 ^ b_l), 16 * (4 ^ b_l), 16 * (5 ^ b_l)};                 // [k4]: q = 0, 1, 4, 5
  const int wrow = row0 + wm * 32;
  for (int it = 0; it < nkt; ++it) {
    if (it + TPF < nkt) issue(it + TPF);
    const int s = it % TST;
    const int k0 = (kt_begin + it) * BK;
    mbar_wait(&full_bar[s], (uint32_t)(it / TST) & 1u);
    const unsigned char* stage = ring + (size_t)s * T_STAGE_BYTES;
    constexpr int Q2[4] = {0, 2, 8, 10};
    if (a_upper && k0 < wrow + 32) {
      // diagonal block of an upper-triangular A: the largest k of step k4 is k0 + Q2[k4] + 5; a group of 8
      // rows that starts below it contributes nothing
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        const int kmax = k0 + Q2[k4] + 5;
        if (wrow > kmax) continue;
        double bf[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) bf[j] = *reinterpret_cast<const double*>(stage + b_lane + j * 1024 + b_off[k4]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (wrow + i * 8 > kmax) continue;
          const double af = *reinterpret_cast<const double*>(stage + a_lane + (i >> 1) * 2048 + Q2[k4] * 128 + a_off[i & 1][k4 & 1]);
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af, bf[j]);
        }
      }
    } else {
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        double af[4], bf[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          af[i] = *reinterpret_cast<const double*>(stage + a_lane + (i >> 1) * 2048 + Q2[k4] * 128 + a_off[i & 1][k4 & 1]);
#pragma unroll
        for (int j = 0; j < 4; ++j) bf[j] = *reinterpret_cast<const double*>(stage + b_lane + j * 1024 + b_off[k4]);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);
  }

  if constexpr (FUSE) {
    // a thread holds, for each of 4 rows (i), 8 values: columns wn 32 + 8 j + 2 fk + {0, 1}; the 64 columns of a
    // row sit in the 4 lanes of a quad (fk) of two warps (wn)
    const int quad0 = lane & ~3;
    bool okc[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int cc = col0 + wn * 32 + j * 8 + 2 * fk;
      okc[2 * j] = cc < n; okc[2 * j + 1] = cc + 1 < n;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = row0 + wm * 32 + i * 8 + fr;
      double rs = 0.0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (okc[2 * j]) rs += acc[i][j][0];
        if (okc[2 * j + 1]) rs += acc[i][j][1];
      }
      rs += __shfl_xor_sync(0xffffffffu, rs, 1);
      rs += __shfl_xor_sync(0xffffffffu, rs, 2);
      if (fk == 0 && r < m) bf.tile_sum[(size_t)(2 * bn + wn) * m + r] = rs;
    }
    const int np = bf.nprobs;
    for (int kq = 0; kq < np; ++kq) {
      unsigned inm[4], off[4], tot[4], nbq[4], base[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {                       // masks and counts of the four rows
        const int r = row0 + wm * 32 + i * 8 + fr;
        const bool rv = r < m;
        double lo = 0.0, hi = 0.0;
        if (rv) {
          const ulonglong2 br = __ldg(reinterpret_cast<const ulonglong2*>(bf.brackets) + (size_t)r * np + kq);
          lo = band_bound(br.x); hi = band_bound(br.y);
        }
        unsigned nb = 0u, mk = 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const double x = acc[i][j][e];
            const bool ok = okc[2 * j + e] && rv;          // rows and columns past the end: nothing below, nothing inside
            nb += (ok && x < lo) ? 1u : 0u;
            mk |= (ok && x >= lo && x <= hi) ? (1u << (2 * j + e)) : 0u;
          }
        nb += __shfl_xor_sync(0xffffffffu, nb, 1);
        nb += __shfl_xor_sync(0xffffffffu, nb, 2);
        const unsigned c = __popc(mk);
        const unsigned c0 = __shfl_sync(0xffffffffu, c, quad0), c1 = __shfl_sync(0xffffffffu, c, quad0 + 1),
                       c2 = __shfl_sync(0xffffffffu, c, quad0 + 2), c3 = __shfl_sync(0xffffffffu, c, quad0 + 3);
        inm[i] = mk; nbq[i] = nb;
        tot[i] = c0 + c1 + c2 + c3;
        off[i] = (fk > 0 ? c0 : 0u) + (fk > 1 ? c1 : 0u) + (fk > 2 ? c2 : 0u);
      }
      // the four cursor atomics of this probability go out back to back: one round trip to L2, not four
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = row0 + wm * 32 + i * 8 + fr;
        base[i] = 0u;
        if (fk == 0 && r < m) {
          if (nbq[i]) atomicAdd(bf.below + (size_t)r * np + kq, nbq[i]);
          if (tot[i]) base[i] = atomicAdd(bf.cursor + (size_t)r * np + kq, tot[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        unsigned p = __shfl_sync(0xffffffffu, base[i], quad0) + off[i];
        if (inm[i] == 0u) continue;
        const int r = row0 + wm * 32 + i * 8 + fr;
        double* seg = bf.buf + ((size_t)r * np + kq) * bf.cap;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e)
            if (inm[i] >> (2 * j + e) & 1u) { if (p < bf.cap) seg[p] = acc[i][j][e]; ++p; }   // an overflowing row is flagged by the selection
      }
    }
    return;
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

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      p = nullptr;
    }
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}
// 2-D column-major FP64 matrix (rows x cols, leading dimension ld) with a box of box_r x box_c, 128-byte swizzle
static bool make_map(CUtensorMap* map, const double* base, int rows, int cols, int64_t ld, int box_r, int box_c) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)rows, (cuuint64_t)cols};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(double)};
  const cuuint32_t box[2] = {(cuuint32_t)box_r, (cuuint32_t)box_c};
  const cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// A[kk, col] = N(0,1) draw of (chain = active[col / attempts], draw index = pos[chain] + col % attempts,
// coordinate kk; with col_off the columns of active chain a are [col_off[a], col_off[a+1])) — the same Philox addressing the chain kernels use, so pool and in-kernel
// draws agree.
__global__ void __launch_bounds__(256) fill_normals_kernel(int d, size_t n_cols, double* __restrict__ A, NormalGen g) {
  // one column per CTA at a time: which chain and draw a column belongs to is worked out once per column
  // (no 64-bit divisions per element), the threads run over the coordinate pairs
  const int pairs = (d + 1) / 2;
  const bool d_even = (d & 1) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0;
  for (size_t col = blockIdx.x; col < n_cols; col += gridDim.x) {
    int a; int64_t j;
    if (g.col_off) {                                   // per-chain column ranges: last a with col_off[a] <= col
      int lo = 0, hi = g.n_active;
      while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if ((size_t)g.col_off[mid] <= col) lo = mid; else hi = mid; }
      a = lo; j = (int64_t)(col - (size_t)g.col_off[lo]);
    } else {
      a = (int)(col / (size_t)g.attempts); j = (int64_t)(col - (size_t)a * g.attempts);
    }
    const int chain = g.active ? g.active[a] : a;
    const uint64_t draw = (uint64_t)(g.pos[chain] + j);
    const uint64_t gchain = (uint64_t)(g.chain_offset + (int64_t)chain * g.chain_stride);
    double* dst = A + col * d;
    for (int pr = threadIdx.x; pr < pairs; pr += blockDim.x) {
      double n0, n1;
      normal_pair(g.seed, gchain, draw, (uint32_t)pr, n0, n1);
      if (d_even) *reinterpret_cast<double2*>(dst + 2 * pr) = make_double2(n0, n1);   // col * d + 2 pr is even: 16-byte aligned
      else { dst[2 * pr] = n0; if (2 * pr + 1 < d) dst[2 * pr + 1] = n1; }
    }
  }
}

cudaError_t launch_dgemm(int m, int n, int k, const double* A, int64_t lda, const double* B,
                         int64_t ldb, double* C, int64_t ldc, double* row_sum, int a_upper,
                         cudaStream_t st, int b_blk_w, int64_t b_blk_pitch) {
  if (m <= 0 || n <= 0) return cudaSuccess;
  // kernel choice (LQG_GEMM_RING: 2 = tensor-map TMA where possible (default), 1 = one-column bulk copies,
  // 0 = register-staged kernel): bulk copies of either kind need 16-byte aligned operand columns; the
  // tensor-map form also needs plain (not blocked) columns of B
  static const char* env_ring = getenv("LQG_GEMM_RING");
  const int want = env_ring ? atoi(env_ring) : 2;
  const bool aligned = ((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0) && lda % 2 == 0 && ldb % 2 == 0 &&
                       (b_blk_w <= 0 || b_blk_pitch % 2 == 0);
  int kind = !aligned ? 0 : std::min(want, b_blk_w > 0 ? 1 : 2);
  CUtensorMap mapA, mapB;
  if (kind == 2 && !(make_map(&mapA, A, m, k, lda, 16, BK) && make_map(&mapB, B, k, n, ldb, BK, BN))) kind = 1;
  const size_t smem = kind == 2 ? (size_t)TST * T_STAGE_BYTES + 1024
                    : kind == 1 ? (size_t)ST * STAGE * sizeof(double) : (size_t)(2 * BK * A_LD + 2 * BN * B_LD) * sizeof(double);
  const void* kern = kind == 2 ? (const void*)dgemm_tma_kernel<false> : kind == 1 ? (const void*)dgemm_ring_kernel : (const void*)dgemm_dmma_kernel;
  cudaError_t e0 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e0 != cudaSuccess) return e0;
  const long long tiles = (long long)((m + BM - 1) / BM) * ((n + BN - 1) / BN);
  if (tiles > 0x7fffffffLL) return cudaErrorInvalidValue;
  int tail_cb = 0;
  if (a_upper) {                       // four waves of resident CTAs' worth of column blocks, longest tiles first
    static int slots_of[3] = {0, 0, 0};
    if (slots_of[kind] == 0) {
      int dev = 0, sms = 148, per_sm = 2;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem);
      slots_of[kind] = sms * (per_sm > 0 ? per_sm : 1);
    }
    static const char* env_tail = getenv("LQG_GEMM_TAIL");           // tuning aid: waves of column blocks reordered (default 4)
    const double waves = env_tail ? atof(env_tail) : 4.0;
    const int MBc = (m + BM - 1) / BM, NBc = (n + BN - 1) / BN;
    tail_cb = std::min(NBc, (int)(waves * slots_of[kind] / MBc));
  }
  if (kind == 2)
    dgemm_tma_kernel<false><<<(unsigned)tiles, NT, smem, st>>>(mapA, mapB, m, n, k, C, ldc, row_sum, a_upper, tail_cb, BandFuse{});
  else if (kind == 1)
    dgemm_ring_kernel<<<(unsigned)tiles, NT, smem, st>>>(m, n, k, A, lda, B, ldb, C, ldc, row_sum, a_upper, b_blk_w, b_blk_pitch, tail_cb);
  else
    dgemm_dmma_kernel<<<(unsigned)tiles, NT, smem, st>>>(m, n, k, A, lda, B, ldb, C, ldc, row_sum, a_upper, b_blk_w, b_blk_pitch, tail_cb);
  ++launch_counter();
  return cudaGetLastError();
}

// the tensor-map kernel needs 16-byte aligned operand columns
bool dgemm_bands_usable(const double* A, int64_t lda, const double* B, int64_t ldb) {
  static const char* env_ring = getenv("LQG_GEMM_RING");
  if (env_ring && atoi(env_ring) != 2) return false;
  return ((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0) && lda % 2 == 0 && ldb % 2 == 0 && encode_tiled_fn() != nullptr;
}

int dgemm_bands_parts(int n) { return 2 * ((n + BN - 1) / BN); }

// A (m x k) B (k x n) summarised for the bands, nothing stored (dgemm_tma_kernel<true>)
cudaError_t launch_dgemm_bands(int m, int n, int k, const double* A, int64_t lda, const double* B, int64_t ldb,
                               const BandFuse& f, cudaStream_t st) {
  if (m <= 0 || n <= 0) return cudaSuccess;
  CUtensorMap mapA, mapB;
  if (!dgemm_bands_usable(A, lda, B, ldb) || !(make_map(&mapA, A, m, k, lda, 16, BK) && make_map(&mapB, B, k, n, ldb, BK, BN)))
    return cudaErrorNotSupported;
  const size_t smem = (size_t)TST * T_STAGE_BYTES + 1024;
  cudaError_t e0 = cudaFuncSetAttribute(dgemm_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e0 != cudaSuccess) return e0;
  const long long tiles = (long long)((m + BM - 1) / BM) * ((n + BN - 1) / BN);
  if (tiles > 0x7fffffffLL) return cudaErrorInvalidValue;
  dgemm_tma_kernel<true><<<(unsigned)tiles, NT, smem, st>>>(mapA, mapB, m, n, k, nullptr, 0, nullptr, 0, 0, f);
  ++launch_counter();
  return cudaGetLastError();
}

cudaError_t launch_fill_normals(int d, size_t n_cols, double* A, const NormalGen& g, cudaStream_t st) {
  fill_normals_kernel<<<(unsigned)std::min<size_t>(n_cols, 148 * 16), 256, 0, st>>>(d, n_cols, A, g);
  ++launch_counter();
  return cudaGetLastError();
}

}  // namespace lqg
