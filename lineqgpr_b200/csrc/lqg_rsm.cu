// Batched rejection sampling from the mode (Maatouk & Bay), R/lineqGPsamplers.R:36-61.
//
// Reference loop, per output sample:
//   repeat { repeat x = mvrnorm(1, map, Sigma) until all(lb <= x <= ub)     (:50-53)
//            t = exp((map - x)' Sigma^-1 map); u = runif(1) } until u <= t   (:54-56)
// Here a ROUND evaluates P proposals at once; the output is the first nsim accepted
// proposals IN PROPOSAL ORDER, so given the same draws the accept/reject decisions and the
// emitted columns are the reference's.  Uniforms are consumed one per
// IN-BOX proposal (:55): with injected draws the uniform index is the in-box rank, obtained
// from an ordered prefix sum; with Philox both streams are addressed by the proposal index.
//
//   rsm_eval   : one warp per proposal; z (Philox / injected) staged in shared memory,
//                x = map + factor z with coalesced column reads, warp-vote box test,
//                shuffle-reduced exponent.  Per block: count of in-box proposals.
//   rsm_scan   : exclusive scan of the per-block counts (single CTA).
//   rsm_accept : in-box rank -> uniform -> accept flag; per-block accept counts.
//   rsm_emit   : ordered compaction: ballot/popc prefix inside the warp, scanned block
//                offsets across CTAs; accepted proposals are recomputed and written to
//                their output column.
#include "lqg_common.cuh"
#include "lqg_kernels.h"
#include "lqg_scan.cuh"

namespace lqg {

namespace {

constexpr int PPB = kScanBlock;   // proposals per block (= threads per block in accept/emit)
constexpr int WPB = 8;        // warps per block in eval/emit
constexpr uint32_t TAG_Z = 0x52534D31u, TAG_U = 0x52534D32u;

__device__ __forceinline__ void rsm_normal_pair(uint64_t seed, uint64_t p, uint32_t pair,
                                                double& n0, double& n1) {
  uint32_t r[4];
  Philox::gen(pair, (uint32_t)p, (uint32_t)(p >> 32), TAG_Z, (uint32_t)seed, (uint32_t)(seed >> 32), r);
  double u1 = u01(r[0], r[1]), u2 = u01(r[2], r[3]);
  double rad = sqrt(-2.0 * log(u1));
  double s, c;
  sincospi(2.0 * u2, &s, &c);
  n0 = rad * c; n1 = rad * s;
}
__device__ __forceinline__ double rsm_uniform(uint64_t seed, uint64_t p) {
  uint32_t r[4];
  Philox::gen(0u, (uint32_t)p, (uint32_t)(p >> 32), TAG_U, (uint32_t)seed, (uint32_t)(seed >> 32), r);
  return u01(r[0], r[1]);
}

// z of proposal p into zs[0..d), by one warp
__device__ __forceinline__ void load_z(const RsmParams& q, int64_t p_local, double* zs, int lane) {
  const int d = q.d;
  if (q.inj_z) {
    const double* src = q.inj_z + (size_t)(q.first_proposal + p_local) * d;
    for (int i = lane; i < d; i += 32) zs[i] = src[i];
  } else {
    const uint64_t gp = (uint64_t)(q.first_proposal + p_local);
    for (int pr = lane; 2 * pr < d; pr += 32) {
      double n0, n1;
      rsm_normal_pair(q.seed, gp, (uint32_t)pr, n0, n1);
      zs[2 * pr] = n0;
      if (2 * pr + 1 < d) zs[2 * pr + 1] = n1;
    }
  }
  __syncwarp();
}

// x_i = map_i + sum_k factor[i,k] z_k   (MASS::mvrnorm: mu + eS$vectors diag(sqrt(ev)) z)
__device__ __forceinline__ double proposal_coord(const RsmParams& q, const double* zs, int i) {
  const int d = q.d;
  double acc = 0.0;
  for (int k = 0; k < d; ++k) acc = fma(__ldg(q.factor + (size_t)k * d + i), zs[k], acc);
  return __ldg(q.map + i) + acc;
}

}  // namespace

__global__ void __launch_bounds__(WPB * 32)
rsm_eval_kernel(const RsmParams q, int* block_inbox) {
  extern __shared__ double zs_all[];
  __shared__ int cnt;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int d = q.d;
  double* zs = zs_all + (size_t)warp * d;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  int my_inbox = 0;
  for (int j = warp; j < PPB; j += WPB) {
    const int64_t p = (int64_t)blockIdx.x * PPB + j;
    if (p >= q.n_proposals) break;
    load_z(q, p, zs, lane);
    bool in = true;
    double dot = 0.0;
    for (int i = lane; i < d; i += 32) {
      const double x = proposal_coord(q, zs, i);
      in = in && (x >= __ldg(q.lb + i)) && (x <= __ldg(q.ub + i));            // :51
      dot = fma(__ldg(q.map + i) - x, __ldg(q.w + i), dot);                    // :54
    }
    in = __all_sync(0xffffffffu, in);
    dot = warp_sum(dot);
    if (lane == 0) {
      q.flags[p] = in ? 1 : 0;
      q.tval[p] = in ? exp(dot) : 0.0;
      my_inbox += in ? 1 : 0;
    }
    __syncwarp();
  }
  if (lane == 0 && my_inbox) atomicAdd(&cnt, my_inbox);
  __syncthreads();
  if (threadIdx.x == 0) block_inbox[blockIdx.x] = cnt;
}

__global__ void __launch_bounds__(PPB)
rsm_accept_kernel(const RsmParams q, const int64_t* inbox_off, int* block_acc) {
  __shared__ int wt[PPB / 32];
  const int64_t p = (int64_t)blockIdx.x * PPB + threadIdx.x;
  const bool valid = p < q.n_proposals;
  const int in = valid ? (q.flags[p] & 1) : 0;
  int tot;
  const int local = block_excl_scan(in, wt, &tot);
  int acc = 0;
  if (in) {
    const int64_t rank = q.inbox_before + inbox_off[blockIdx.x] + local;      // uniforms used so far
    double u;
    bool have = true;
    if (q.inj_u) { have = rank < q.n_inj_u; u = have ? q.inj_u[rank] : 2.0; }
    else u = rsm_uniform(q.seed, (uint64_t)(q.first_proposal + p));           // :55
    acc = (have && u <= q.tval[p]) ? 1 : 0;                                   // :56
    q.flags[p] = (unsigned char)(1 | (acc << 1) | (have ? 0 : 4));
  }
  const int nacc = __syncthreads_count(acc);
  if (threadIdx.x == 0) block_acc[blockIdx.x] = nacc;
}

// counters: [0] proposals consumed up to the nsim-th accept, [1] in-box up to there
__global__ void __launch_bounds__(PPB)
rsm_emit_kernel(const RsmParams q, const int64_t* acc_off, const int64_t* inbox_off,
                int64_t slot_base, int64_t nsim, double* out, unsigned long long* counters) {
  extern __shared__ double zs_all[];
  __shared__ int wt[PPB / 32];
  __shared__ int list_p[PPB];
  __shared__ long long list_slot[PPB];
  __shared__ int n_list;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int d = q.d;
  const int64_t p = (int64_t)blockIdx.x * PPB + threadIdx.x;
  const bool valid = p < q.n_proposals;
  const int fl = valid ? q.flags[p] : 0;
  const int acc = (fl >> 1) & 1;
  int tot;
  // ordered compaction: ballot/popc inside block_excl_scan's warp stage, block offsets across CTAs
  const int local = block_excl_scan(acc, wt, &tot);
  if (threadIdx.x == 0) n_list = tot;
  if (acc) {
    const int64_t slot = slot_base + acc_off[blockIdx.x] + local;
    list_p[local] = threadIdx.x;
    list_slot[local] = slot;
    if (slot == nsim - 1)             // the reference would stop right here
      counters[0] = (unsigned long long)(q.first_proposal + p + 1);
  }
  __syncthreads();
  // in-box count up to the nsim-th accept: recomputed by the owning block
  {
    const int in = fl & 1;
    int tin;
    const int lin = block_excl_scan(in, wt, &tin);
    if (acc && slot_base + acc_off[blockIdx.x] + local == nsim - 1)
      counters[1] = (unsigned long long)(q.inbox_before + inbox_off[blockIdx.x] + lin + 1);
  }
  double* zs = zs_all + (size_t)warp * d;
  for (int e = warp; e < n_list; e += PPB / 32) {
    const long long slot = list_slot[e];
    if (slot >= nsim) continue;
    const int64_t pl = (int64_t)blockIdx.x * PPB + list_p[e];
    load_z(q, pl, zs, lane);
    double* dst = out + (size_t)slot * d;
    for (int i = lane; i < d; i += 32) dst[i] = proposal_coord(q, zs, i);     // :58
    __syncwarp();
  }
}

int rsm_blocks(int64_t n_proposals) { return (int)((n_proposals + PPB - 1) / PPB); }

cudaError_t launch_rsm_round(const RsmParams& q, int* block_cnt, int64_t* inbox_off, int64_t* acc_off,
                             int64_t slot_base, int64_t nsim, double* out,
                             unsigned long long* counters, cudaStream_t st) {
  const int nb = rsm_blocks(q.n_proposals);
  const size_t smem = (size_t)WPB * q.d * sizeof(double);
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(rsm_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(rsm_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
  rsm_eval_kernel<<<nb, WPB * 32, smem, st>>>(q, block_cnt);
  scan_counts_kernel<<<1, 1024, 0, st>>>(block_cnt, inbox_off, nb);
  rsm_accept_kernel<<<nb, PPB, 0, st>>>(q, inbox_off, block_cnt);
  scan_counts_kernel<<<1, 1024, 0, st>>>(block_cnt, acc_off, nb);
  rsm_emit_kernel<<<nb, PPB, smem, st>>>(q, acc_off, inbox_off, slot_base, nsim, out, counters);
  launch_counter() += 5;
  return cudaGetLastError();
}

}  // namespace lqg
