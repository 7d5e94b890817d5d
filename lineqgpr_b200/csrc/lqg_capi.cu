// C-ABI layer of lineqgpr_b200 (see include/lineqgpr_b200.h): argument validation, staging
// of host buffers, launch orchestration, status codes.  No compute happens on the host.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <atomic>
#include <cmath>
#include <mutex>
#include <vector>

#include <exception>
#include <functional>
#include <memory>

#include "lqg_comm.h"
#include "lqg_common.cuh"
#include "lqg_hostcopy.h"
#include "lqg_kernels.h"

namespace lqg {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static std::atomic<int64_t> g_launches{0};
static thread_local int64_t g_launch_tls = 0;
int64_t& launch_counter() {
  // kernels are launched from the calling thread; fold the thread-local count into the
  // global one lazily (see lqg_launch_count)
  return g_launch_tls;
}
static void fold_launches() {
  g_launches += g_launch_tls;
  g_launch_tls = 0;
}

// owning device allocation from the stream-ordered pool: after the first call the pool keeps
// the memory (release threshold = max), so repeated calls do not pay cudaMalloc/cudaFree
static void retain_pool_once() {
  static std::once_flag once;
  std::call_once(once, [] {
    int dev = 0;
    cudaMemPool_t pool;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t thr = UINT64_MAX;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
  });
}
// owning stream / event handles: released on every return path
struct StreamGuard {
  cudaStream_t s = nullptr;
  ~StreamGuard() { if (s) cudaStreamDestroy(s); }
};
struct EventGuard {
  cudaEvent_t e = nullptr;
  ~EventGuard() { if (e) cudaEventDestroy(e); }
};
// LQG_GUARD=1 (debugging aid): every device buffer gets a 256-byte canary band on either side,
// checked when the buffer is released; lqg_guard_violations() reports how many were overwritten.
// This is the library's own out-of-bounds-write detector for small cases.
static std::atomic<int> g_guard_violations{0};
static bool guard_mode() {
  static const bool on = [] { const char* e = getenv("LQG_GUARD"); return e && *e && *e != '0'; }();
  return on;
}
struct DevBuf {
  static constexpr size_t kGuard = 256;
  static constexpr int kPattern = 0xA5;
  void* p = nullptr;
  size_t bytes = 0;
  cudaStream_t st = nullptr;
  static thread_local cudaStream_t cur_stream;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (!p) return;
    if (!guard_mode()) { cudaFreeAsync(p, st); p = nullptr; return; }
    unsigned char* base = static_cast<unsigned char*>(p) - kGuard;
    unsigned char h[2 * kGuard];
    bool ok = cudaStreamSynchronize(st) == cudaSuccess &&
              cudaMemcpy(h, base, kGuard, cudaMemcpyDeviceToHost) == cudaSuccess &&
              cudaMemcpy(h + kGuard, base + kGuard + bytes, kGuard, cudaMemcpyDeviceToHost) == cudaSuccess;
    if (ok) {
      for (size_t i = 0; i < 2 * kGuard; ++i)
        if (h[i] != kPattern) {
          ++g_guard_violations;
          fprintf(stderr, "[lqg guard] %s band of a %zu-byte buffer overwritten at offset %zu\n",
                  i < kGuard ? "leading" : "trailing", bytes, i < kGuard ? i : i - kGuard);
          break;
        }
    } else {
      cudaGetLastError();
    }
    cudaFreeAsync(base, st);
    p = nullptr;
  }
  cudaError_t alloc(size_t n) {
    release();
    retain_pool_once();
    bytes = n ? n : 16;
    st = cur_stream;
    if (!guard_mode()) return cudaMallocAsync(&p, bytes, st);
    bytes = (bytes + 15) & ~size_t(15);
    void* base = nullptr;
    cudaError_t e = cudaMallocAsync(&base, bytes + 2 * kGuard, st);
    if (e != cudaSuccess) return e;
    unsigned char* b = static_cast<unsigned char*>(base);
    if ((e = cudaMemsetAsync(b, kPattern, kGuard, st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(b + kGuard + bytes, kPattern, kGuard, st)) != cudaSuccess) return e;
    p = b + kGuard;
    return cudaSuccess;
  }
  template <class T> T* as() { return reinterpret_cast<T*>(p); }
};
thread_local cudaStream_t DevBuf::cur_stream = nullptr;

// A read-only input that is either already on the device or staged from the host.
struct Input {
  DevBuf buf;
  const double* dev = nullptr;
  int stage(const double* src, size_t n, int mem, cudaStream_t st, size_t pad_to = 0) {
    if (mem == LQG_MEM_DEVICE && pad_to == 0) { dev = src; return LQG_OK; }
    const size_t n_alloc = pad_to > n ? pad_to : n;
    LQG_CUDA_TRY(buf.alloc(n_alloc * sizeof(double)));
    if (n_alloc > n) LQG_CUDA_TRY(cudaMemsetAsync(buf.p, 0, n_alloc * sizeof(double), st));
    LQG_CUDA_TRY(cudaMemcpyAsync(buf.p, src, n * sizeof(double),
                                 mem == LQG_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    dev = buf.as<double>();
    return LQG_OK;
  }
};

// Device -> host copy of a large result: a pageable destination (R / numpy matrix) goes through the
// pinned staging ring and helper threads (lqg_hostcopy.cu), anything else through the copy engine.
// Ordered after the work already enqueued on st; returns after the bytes are in place when staged.
static cudaError_t d2h_result(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t rows,
                              cudaStream_t st) {
  static const char* env_staged = getenv("LQG_STAGED");
  if (width * rows >= ((size_t)32 << 20) && !(env_staged && env_staged[0] == '0') && StagedCopier::is_pageable(dst)) {
    StagedCopier sc(st);
    cudaError_t e = sc.copy2d(dst, dpitch, src, spitch, width, rows);
    if (e != cudaSuccess) return e;
    return sc.finish();
  }
  if (rows == 1 || (dpitch == width && spitch == width))
    return cudaMemcpyAsync(dst, src, width * rows, cudaMemcpyDeviceToHost, st);
  return cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, cudaMemcpyDeviceToHost, st);
}

static int have_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    set_error("no CUDA device available (%s); lineqgpr_b200 has no CPU fallback",
              e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return 0;
  }
  return n;
}

static int fetch_host(std::vector<double>& dst, const double* src, size_t n, int mem, cudaStream_t st) {
  dst.resize(n);
  if (mem == LQG_MEM_DEVICE) {
    LQG_CUDA_TRY(cudaMemcpyAsync(dst.data(), src, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    LQG_CUDA_TRY(cudaStreamSynchronize(st));
  } else {
    memcpy(dst.data(), src, n * sizeof(double));
  }
  return LQG_OK;
}

static void fill_root_stats(lqg_hmc_stats* s, const unsigned long long* h, double chain_ms, double pre_ms) {
  if (!s) return;
  s->root_transitions = (int64_t)h[0]; s->root_bounces = (int64_t)h[1]; s->root_hit_searches = (int64_t)h[2];
  s->root_restarts = (int64_t)(h[3] + h[4]); s->root_chain_ms = chain_ms; s->root_prepass_ms = pre_ms;
}
static void fill_stats(lqg_hmc_stats* s, const unsigned long long* h, double chain_ms, double pre_ms, unsigned long long slow = 0,
                       unsigned long long widen = 0) {
  if (!s) return;
  s->slow_searches = (int64_t)slow;
  s->window_retries = (int64_t)widen;
  s->transitions = (int64_t)h[0]; s->bounces = (int64_t)h[1]; s->hit_searches = (int64_t)h[2];
  s->vel_restarts = (int64_t)h[3]; s->chk_restarts = (int64_t)h[4]; s->exact_evals = (int64_t)h[5];
  s->violations = (int64_t)h[6]; s->failed_chains = (int64_t)h[7];
  s->chain_ms = chain_ms; s->prepass_ms = pre_ms;
}

// LQG_TRACE=1 prints host wall-clock phase timings of the entry points to stderr
struct Phase {
  bool on; double t0; const char* who;
  static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
  explicit Phase(const char* w) : on(getenv("LQG_TRACE") != nullptr), t0(now()), who(w) {}
  void mark(const char* what) { if (on) { double t = now(); fprintf(stderr, "[lqg %s] %-18s %8.2f ms\n", who, what, t - t0); t0 = t; } }
};

struct Timer {
  cudaEvent_t a = nullptr, b = nullptr;
  cudaStream_t st;
  explicit Timer(cudaStream_t s) : st(s) { cudaEventCreate(&a); cudaEventCreate(&b); }
  ~Timer() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
  void start() { cudaEventRecord(a, st); }
  void stop() { cudaEventRecord(b, st); }
  double ms() { float f = 0; cudaEventSynchronize(b); cudaEventElapsedTime(&f, a, b); return f; }
};

static lqg_opts default_opts() {
  lqg_opts o;
  memset(&o, 0, sizeof(o));
  o.mem = LQG_MEM_HOST; o.n_chains = 1; o.prepass = -1;
  return o;
}

// summarise per-chain status words into a return code
static int chain_status_rc(const std::vector<int>& st, unsigned long long* failed) {
  int rc = LQG_OK;
  unsigned long long f = 0;
  for (int v : st) if (v != 0) { ++f; if (rc == LQG_OK) rc = v; }
  *failed = f;
  return rc;
}

}  // namespace lqg

using namespace lqg;

// No exception crosses the C ABI: host-side allocation failures (std::bad_alloc), thread start-up
// failures (std::system_error) and anything else thrown inside an entry point become a status.
#define LQG_API_BEGIN try {
#define LQG_API_END(name)                                                      \
  } catch (const std::exception& e) {                                          \
    set_error("%s: internal error: %s", name, e.what());                      \
    return LQG_ERR_INTERNAL;                                                   \
  } catch (...) {                                                              \
    set_error("%s: internal error (unknown exception)", name);                \
    return LQG_ERR_INTERNAL;                                                   \
  }

static int ul_factor_device(int d, const double* S_dev, double* U_dev, cudaStream_t st);

// ===========================================================================================
// Box-form HMC with every operand resident on the device.  lqg_hmc_box (host or device pointers)
// and lqg_simulate (which keeps eta on the device and feeds it to the xi / path stages) both end
// up here.
struct BoxRun {
  int d = 0;
  const double* mu = nullptr; const double* U = nullptr; int u_upper = 1;
  const double* Sigma = nullptr; const double* lb = nullptr; const double* ub = nullptr;
  const double* init = nullptr; int64_t init_ld = 0;
  int n_chains = 1, n_per_chain = 0, burn_in = 0;
  uint64_t seed = 0; int64_t chain_offset = 0; int prepass = -1; int max_restarts = 0;
  int exact_search = 0;                                            // 1: every hit search with tmg's own arithmetic
  int no_window = 0;                                               // 1: fast search over the whole remaining time
  const double* injected = nullptr; int64_t n_injected = 0;      // device pointer
  double* out = nullptr;                                          // d x (n_chains*n_per_chain), chain-major columns
  cudaStream_t st = nullptr;
  // Called (host side) whenever samples [from, upto) of EVERY chain are final in `out`; `ready`
  // has been recorded on st after the kernels that wrote them.  Lets a consumer (D2H copy, xi GEMM,
  // all-gather) work on another stream while later rounds run.  Nullable.
  std::function<int(int from, int upto, cudaEvent_t ready)> on_round;
  bool short_tail = false;    // cut the last stretch into halving rounds (the consumer cannot overlap the last one)
  int fork = 0, fork_burn_in = 0;   // lqg_opts.fork: families of `fork` chains share one root chain's burn-in
  // set by the fork logic for its two phases (never by the entry points):
  int chain_stride = 1;             // global id of local chain c = chain_offset + c * chain_stride
  const double* state0 = nullptr;   // centred start states d x n_chains (instead of init - mu)
  const int64_t* pos0 = nullptr;    // momenta already consumed per chain
  const int* sdone0 = nullptr;      // transitions already behind every chain
  bool keep_pos = false;            // leave the chains' draw counters in BoxResult::pos
};
struct BoxResult {
  unsigned long long stats[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // device layout, see BoxParams::stats
  std::vector<int> status;
  double chain_ms = 0.0, pre_ms = 0.0;
  DevBuf state;               // d x n_chains centred positions x_b of the last transition
  DevBuf pos;                 // [n_chains] momenta consumed (BoxRun::keep_pos)
  int max_restarts = 0;
  unsigned long long root_stats[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // the root phase's share of stats (forked families)
  double root_chain_ms = 0.0, root_pre_ms = 0.0;
};

static int hmc_box_rounds(const BoxRun& r, BoxResult& res) {
  Phase ph("hmc_box");
  const int d = r.d, n_chains = r.n_chains, n_per_chain = r.n_per_chain, burn_in = r.burn_in;
  cudaStream_t st = r.st;
  const size_t n_cols = (size_t)n_chains * n_per_chain;
  double* d_out = r.out;
  DevBuf b_glo, b_ghi, b_sd, b_misc, b_status, b_sdone, b_rcount, b_act;
  DevBuf& b_state = res.state;
  DevBuf& b_pos = res.pos;
  LQG_CUDA_TRY(b_state.alloc((size_t)d * n_chains * sizeof(double)));
  LQG_CUDA_TRY(b_glo.alloc(d * sizeof(double)));
  LQG_CUDA_TRY(b_ghi.alloc(d * sizeof(double)));
  LQG_CUDA_TRY(b_sd.alloc(d * sizeof(double)));
  LQG_CUDA_TRY(b_misc.alloc(16 * sizeof(unsigned long long)));      // [0..7] stats, [8] queue head, [9] n_active
  LQG_CUDA_TRY(b_status.alloc((size_t)n_chains * sizeof(int)));
  LQG_CUDA_TRY(b_sdone.alloc((size_t)n_chains * sizeof(int)));
  LQG_CUDA_TRY(b_pos.alloc((size_t)n_chains * sizeof(int64_t)));
  LQG_CUDA_TRY(b_rcount.alloc((size_t)n_chains * sizeof(int)));
  LQG_CUDA_TRY(b_act.alloc((size_t)2 * n_chains * sizeof(int)));
  LQG_CUDA_TRY(cudaMemsetAsync(b_misc.p, 0, b_misc.bytes, st));
  LQG_CUDA_TRY(cudaMemsetAsync(b_status.p, 0, b_status.bytes, st));
  LQG_CUDA_TRY(cudaMemsetAsync(b_sdone.p, 0, b_sdone.bytes, st));
  LQG_CUDA_TRY(cudaMemsetAsync(b_pos.p, 0, b_pos.bytes, st));
  LQG_CUDA_TRY(cudaMemsetAsync(b_rcount.p, 0, b_rcount.bytes, st));

  ph.mark("alloc");
  LQG_CUDA_TRY(launch_box_prep(d, r.mu, r.lb, r.ub, r.Sigma, b_glo.as<double>(), b_ghi.as<double>(), b_sd.as<double>(), st));
  if (r.state0) {     // forked chains: start states, draw counters and transition counts come from the root phase
    LQG_CUDA_TRY(cudaMemcpyAsync(b_state.p, r.state0, (size_t)d * n_chains * sizeof(double), cudaMemcpyDeviceToDevice, st));
    LQG_CUDA_TRY(cudaMemcpyAsync(b_pos.p, r.pos0, (size_t)n_chains * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    LQG_CUDA_TRY(cudaMemcpyAsync(b_sdone.p, r.sdone0, (size_t)n_chains * sizeof(int), cudaMemcpyDeviceToDevice, st));
  } else {
    // x_b(0) = init - mu for every chain (init_ld == 0 broadcasts one shared start)
    LQG_CUDA_TRY(launch_box_init_state(d, n_chains, r.init, r.init_ld, r.mu, b_state.as<double>(), st));
  }
  std::vector<double> h_sd(d);
  LQG_CUDA_TRY(cudaMemcpyAsync(h_sd.data(), b_sd.p, d * sizeof(double), cudaMemcpyDeviceToHost, st));
  LQG_CUDA_TRY(cudaStreamSynchronize(st));
  double sd_max = 0.0;
  for (int i = 0; i < d; ++i) {
    if (!(h_sd[i] > 0.0)) { set_error("lqg_hmc_box: Sigma[%d,%d] = %g is not positive", i, i, h_sd[i]); return LQG_ERR_INVALID; }
    sd_max = std::max(sd_max, h_sd[i]);
  }

  BoxParams p;
  memset(&p, 0, sizeof(p));
  p.d = d; p.mu = r.mu; p.lb = r.lb; p.ub = r.ub; p.U = r.U; p.u_upper = r.u_upper;
  p.Sigma = r.Sigma; p.glo = b_glo.as<double>(); p.ghi = b_ghi.as<double>(); p.sdiag = b_sd.as<double>();
  p.abs_eps = kFilterRel * std::sqrt(sd_max);
  p.state = b_state.as<double>(); p.n_chains = n_chains;
  p.s_done = b_sdone.as<int>(); p.pos = b_pos.as<int64_t>(); p.rcount = b_rcount.as<int>();
  p.burn_in = burn_in; p.n_per_chain = n_per_chain; p.out = d_out; p.seed = r.seed;
  p.chain_offset = r.chain_offset; p.chain_stride = r.chain_stride; p.injected = r.injected;
  p.n_injected = r.n_injected;
  p.max_restarts = r.max_restarts > 0 ? r.max_restarts : 100000;
  res.max_restarts = p.max_restarts;
  p.stats = b_misc.as<unsigned long long>();
  p.next_chain = reinterpret_cast<int*>(b_misc.as<unsigned long long>() + 8);
  int* d_nactive = reinterpret_cast<int*>(b_misc.as<unsigned long long>() + 9);
  p.chain_status = b_status.as<int>();
  static const char* env_fast = getenv("LQG_BOX_FAST");             // tuning / A-B aid: 0 = exact search only
  p.fast = (r.exact_search || (env_fast && env_fast[0] == '0')) ? 0 : 1;
  static const char* env_win = getenv("LQG_BOX_WINDOW");            // tuning / A-B aid: 0 = always search the whole remaining time
  p.window = (r.no_window || (env_win && env_win[0] == '0')) ? 0 : 1;

  ph.mark("prep");
  const int total = burn_in + n_per_chain;
  // ---- momentum pool: X_a = U A for the next `attempts` draws of every active chain, ONE GEMM ----
  const bool use_pool = !r.injected && (r.prepass == 1 || (r.prepass == -1 && d >= 256));
  double chain_ms = 0.0, pre_ms = 0.0;
  Timer t_chain(st), t_pre(st);
  EventGuard round_guard;
  if (r.on_round) LQG_CUDA_TRY(cudaEventCreateWithFlags(&round_guard.e, cudaEventDisableTiming));
  cudaEvent_t ev_round = round_guard.e;
  int handed = 0;                        // samples per chain already handed to on_round
  int rc;
  if (!use_pool) {
    p.active = nullptr; p.n_active = n_chains; p.attempts = 0x7fffffff; p.pool = nullptr;
    t_chain.start();
    LQG_CUDA_TRY(launch_hmc_box(p, 0, st));
    t_chain.stop();
    chain_ms = t_chain.ms();
  } else {
    // attempts per round: pool + draw buffers stay below ~2 GB each
    const size_t budget = (size_t)2 << 30;
    static const char* env_att = getenv("LQG_POOL_ATTEMPTS");          // tuning aid
    const size_t att_cap = env_att ? (size_t)std::max(4, atoi(env_att)) : 96;   // cfg4: 64 -> 46.3 ms per step, 96 -> 45.4, 128 -> 45.4 (coarser copy-out)
    int attempts = (int)std::max<size_t>(4, std::min<size_t>(att_cap, budget / ((size_t)d * n_chains * sizeof(double))));
    attempts = std::min(attempts, total + 8);
    // the buffers hold max_attempts columns per chain; later rounds never ask for more (the
    // short tail below wants at least 6, hence the floor here rather than there)
    const int max_attempts = std::max(attempts, 6);
    DevBuf b_pool, b_a;
    LQG_CUDA_TRY(b_pool.alloc((size_t)d * n_chains * max_attempts * sizeof(double)));
    LQG_CUDA_TRY(b_a.alloc((size_t)d * n_chains * max_attempts * sizeof(double)));
    ph.mark("pool alloc");
    int n_active = n_chains;
    int* act[2] = {b_act.as<int>(), b_act.as<int>() + n_chains};
    int cur = 0;
    bool first = true;
    // round counters come back through mapped pinned memory (allocated once per thread)
    static thread_local volatile int* h_flags = nullptr;
    static thread_local int* d_flags = nullptr;
    if (!h_flags) {
      void* hp = nullptr;
      LQG_CUDA_TRY(cudaHostAlloc(&hp, 64, cudaHostAllocMapped));
      LQG_CUDA_TRY(cudaHostGetDevicePointer((void**)&d_flags, hp, 0));
      h_flags = (volatile int*)hp;
    }
    // every round hands each active chain at least 4 momenta, so this terminates after at most
    // (total + restarts) / 4 rounds; max_restarts bounds the restarts of every transition
    DevBuf b_off;
    LQG_CUDA_TRY(b_off.alloc((size_t)(n_chains + 1) * sizeof(int)));
    size_t round_cols = (size_t)n_chains * attempts;            // first round: `attempts` columns for every chain
    while (n_active > 0) {
      NormalGen gen{r.seed, r.chain_offset, attempts, first ? nullptr : act[cur], p.pos, first ? nullptr : b_off.as<int>(), n_active, r.chain_stride};
      t_pre.start();
      LQG_CUDA_TRY(launch_fill_normals(d, round_cols, b_a.as<double>(), gen, st));
      LQG_CUDA_TRY(launch_dgemm(d, (int)round_cols, d, r.U, d, b_a.as<double>(), d,
                                b_pool.as<double>(), d, nullptr, r.u_upper, st));
      t_pre.stop();
      p.active = first ? nullptr : act[cur]; p.n_active = n_active; p.attempts = attempts;
      p.col_off = first ? nullptr : b_off.as<int>();
      p.pool = b_pool.as<double>();
      // queue head = 0, n_active = 0, min s_done = 0x7f7f7f7f  (memsets: no copy-engine traffic)
      LQG_CUDA_TRY(cudaMemsetAsync(p.next_chain, 0, 3 * sizeof(int), st));
      LQG_CUDA_TRY(cudaMemsetAsync(d_nactive + 1, 0x7f, sizeof(int), st));
      t_chain.start();
      LQG_CUDA_TRY(launch_hmc_box(p, 0, st));
      t_chain.stop();
      // unfinished chains -> next round's work list, and its plan: pool columns per chain (see box_plan_kernel)
      LQG_CUDA_TRY(launch_box_compact(n_chains, total, p.s_done, p.chain_status, act[cur ^ 1], d_nactive, d_flags,
                                      max_attempts, r.short_tail ? 1 : 0, b_off.as<int>(), st));
      if (ev_round) LQG_CUDA_TRY(cudaEventRecord(ev_round, st));
      LQG_CUDA_TRY(cudaStreamSynchronize(st));
      n_active = h_flags[0];
      const int s_min = std::min((int)h_flags[1], total);
      round_cols = (size_t)h_flags[2];
      attempts = h_flags[3];
      if (r.on_round && s_min - burn_in > handed) {
        const int upto = s_min - burn_in;
        if ((rc = r.on_round(handed, upto, ev_round))) return rc;
        handed = upto;
      }
      pre_ms += t_pre.ms();
      chain_ms += t_chain.ms();
      cur ^= 1;
      first = false;
      ph.mark("round");
    }
  }
  // chains that failed never advance s_min, and the one-launch path has no rounds: whatever is left
  // is handed over here
  if (r.on_round && handed < n_per_chain) {
    LQG_CUDA_TRY(cudaEventRecord(ev_round, st));
    if ((rc = r.on_round(handed, n_per_chain, ev_round))) return rc;
  }

  // ---- audit, stats ----
  LQG_CUDA_TRY(launch_box_violations(d, n_cols, d_out, r.lb, r.ub, p.stats + 6, st));
  res.status.assign(n_chains, 0);
  LQG_CUDA_TRY(cudaMemcpyAsync(res.stats, p.stats, sizeof(res.stats), cudaMemcpyDeviceToHost, st));
  LQG_CUDA_TRY(cudaMemcpyAsync(res.status.data(), b_status.p, n_chains * sizeof(int), cudaMemcpyDeviceToHost, st));
  LQG_CUDA_TRY(cudaStreamSynchronize(st));
  res.chain_ms = chain_ms; res.pre_ms = pre_ms;
  ph.mark("audit");
  return LQG_OK;
}

// Families of chains that share their burn-in (lqg_opts.fork = F > 1): phase 1 runs the root chains
// (one per family, Philox stream of the family's first chain) through the burn-in, phase 2 starts every
// chain of a family from its root's state, fork_burn_in transitions before its first emission.  Which
// root a chain belongs to follows from its GLOBAL id, so a rank whose range starts inside a family
// simply re-runs that family's root: no exchange, and the samples do not depend on the sharding.
static int hmc_box_core(const BoxRun& r, BoxResult& res) {
  const int F = r.fork;
  if (F <= 1 || r.injected || r.burn_in <= 0 || r.init_ld != 0 || r.chain_offset < 0)
    return hmc_box_rounds(r, res);
  const int d = r.d, n = r.n_chains;
  const int64_t g0 = r.chain_offset, g1 = g0 + n - 1;
  const int64_t root0 = g0 - g0 % F;
  const int n_root = (int)((g1 - g1 % F - root0) / F) + 1;
  BoxRun r1 = r;
  r1.fork = 0; r1.n_chains = n_root; r1.chain_offset = root0; r1.chain_stride = F; r1.n_per_chain = 0; r1.out = nullptr;
  r1.on_round = nullptr; r1.short_tail = false; r1.keep_pos = true;
  BoxResult res1;
  int rc;
  if ((rc = hmc_box_rounds(r1, res1))) return rc;
  bool root_failed = false;
  for (int s : res1.status) root_failed |= s != 0;
  if (root_failed) {                     // report through the chains of the failed families
    res.status.assign(n, 0);
    unsigned long long bad = 0;
    for (int c = 0; c < n; ++c) {
      const int s = res1.status[(size_t)((g0 + c - root0) / F)];
      res.status[c] = s; bad += s != 0;
    }
    memcpy(res.stats, res1.stats, sizeof(res.stats));
    res.stats[7] = bad;
    res.chain_ms = res1.chain_ms; res.pre_ms = res1.pre_ms; res.max_restarts = res1.max_restarts;
    return LQG_OK;
  }
  DevBuf b_state0, b_pos0, b_s0;
  LQG_CUDA_TRY(b_state0.alloc((size_t)d * n * sizeof(double)));
  LQG_CUDA_TRY(b_pos0.alloc((size_t)n * sizeof(int64_t)));
  LQG_CUDA_TRY(b_s0.alloc((size_t)n * sizeof(int)));
  const int s0 = std::max(0, r.burn_in - std::max(0, r.fork_burn_in));
  LQG_CUDA_TRY(launch_box_fork(d, n, F, g0, root0, res1.state.as<double>(), res1.pos.as<int64_t>(), b_state0.as<double>(),
                               b_pos0.as<int64_t>(), b_s0.as<int>(), s0, r.st));
  BoxRun r2 = r;
  r2.fork = 0; r2.state0 = b_state0.as<double>(); r2.pos0 = b_pos0.as<int64_t>(); r2.sdone0 = b_s0.as<int>();
  if ((rc = hmc_box_rounds(r2, res))) return rc;
  for (int k : {0, 1, 2, 3, 4, 5, 10, 11}) res.stats[k] += res1.stats[k];
  res.chain_ms += res1.chain_ms; res.pre_ms += res1.pre_ms;
  memcpy(res.root_stats, res1.stats, sizeof(res.root_stats));
  res.root_chain_ms = res1.chain_ms; res.root_pre_ms = res1.pre_ms;
  return LQG_OK;
}

// status of a finished run -> return code + message (shared by lqg_hmc_box and lqg_simulate)
static int box_result_rc(BoxResult& res, const char* who) {
  const int rc = chain_status_rc(res.status, &res.stats[7]);
  if (rc == LQG_ERR_RESTART_CAP) set_error("%s: %llu chain(s) exceeded max_restarts=%d in one transition", who, res.stats[7], res.max_restarts);
  if (rc == LQG_ERR_DRAWS_EXHAUSTED) set_error("%s: injected Gaussian stream too short for %llu chain(s)", who, res.stats[7]);
  if (rc == LQG_OK && res.stats[6] != 0) { set_error("%s: %llu emitted values violate the bounds", who, res.stats[6]); }
  return rc;
}

// host-side validation of the box problem, as the R layers do it
static int validate_box(int d, const std::vector<double>& h_lb, const std::vector<double>& h_ub,
                        const std::vector<double>& h_init, int64_t init_ld, int n_chains) {
  for (int i = 0; i < d; ++i)
    if (!(h_lb[i] < h_ub[i])) {                            // R/lineqGPutils.R:359-360
      set_error("The elements from \"u\" has to be greater than the elements from \"l\" (i=%d)", i);
      return LQG_ERR_INVALID;
    }
  for (int ch = 0; ch < (init_ld == 0 ? 1 : n_chains); ++ch)
    for (int i = 0; i < d; ++i) {
      const double x = h_init[(size_t)ch * init_ld + i];
      if (!(x - h_lb[i] > 0.0) || !(h_ub[i] - x > 0.0)) {  // tmg:R/tmg.R:44-47
        set_error("Error: initial point violates linear constraints. (chain %d, coordinate %d)", ch, i);
        return LQG_ERR_INIT_INFEASIBLE;
      }
    }
  return LQG_OK;
}

extern "C" int lqg_hmc_box(int32_t d, const double* mu, const double* U, int32_t u_upper,
                           const double* Sigma, const double* lb, const double* ub,
                           const double* init, int64_t init_ld, int32_t n_per_chain,
                           int32_t burn_in, const lqg_opts* opts_, double* out,
                           double* state_out, lqg_hmc_stats* stats) {
  LQG_API_BEGIN
  if (stats) memset(stats, 0, sizeof(*stats));
  if (!have_device()) return LQG_ERR_CUDA;
  Phase ph("hmc_box api");
  const lqg_opts o = opts_ ? *opts_ : default_opts();
  const int n_chains = o.n_chains > 0 ? o.n_chains : 1;
  if (d <= 0 || d > 8192 || !mu || !Sigma || !lb || !ub || !init || !out || n_per_chain <= 0 ||
      burn_in < 0 || (init_ld != 0 && init_ld < d)) {
    set_error("lqg_hmc_box: invalid argument (d=%d, n_per_chain=%d, burn_in=%d, init_ld=%lld)", d,
              n_per_chain, burn_in, (long long)init_ld);
    return LQG_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)o.stream;
  DevBuf::cur_stream = st;
  const int mem = o.mem;
  const size_t dd = (size_t)d * d;
  const size_t n_init = init_ld == 0 ? (size_t)d : (size_t)init_ld * (n_chains - 1) + d;

  std::vector<double> h_lb, h_ub, h_init;
  int rc;
  if ((rc = fetch_host(h_lb, lb, d, mem, st)) || (rc = fetch_host(h_ub, ub, d, mem, st)) ||
      (rc = fetch_host(h_init, init, n_init, mem, st)))
    return rc;
  if ((rc = validate_box(d, h_lb, h_ub, h_init, init_ld, n_chains))) return rc;

  ph.mark("validate");
  // ---- stage inputs ----
  Input i_mu, i_U, i_S, i_lb, i_ub, i_init, i_inj;
  if ((rc = i_mu.stage(mu, d, mem, st)) || (U && (rc = i_U.stage(U, dd, mem, st))) ||
      (rc = i_S.stage(Sigma, dd, mem, st)) || (rc = i_lb.stage(lb, d, mem, st)) ||
      (rc = i_ub.stage(ub, d, mem, st)) || (rc = i_init.stage(init, n_init, mem, st)))
    return rc;
  if (o.injected && (rc = i_inj.stage(o.injected, (size_t)n_chains * o.n_injected, mem, st))) return rc;

  if (!U) {                                  // whiten on the device: U U' = Sigma (see lqg_whiten_box)
    LQG_CUDA_TRY(i_U.buf.alloc(dd * sizeof(double)));
    if ((rc = ul_factor_device(d, i_S.dev, i_U.buf.as<double>(), st))) return rc;
    i_U.dev = i_U.buf.as<double>();
    u_upper = 1;
    ph.mark("whiten");
  }

  const size_t n_cols = (size_t)n_chains * n_per_chain;
  DevBuf b_out;
  double* d_out = out;
  if (mem == LQG_MEM_HOST) { LQG_CUDA_TRY(b_out.alloc(n_cols * d * sizeof(double))); d_out = b_out.as<double>(); }
  ph.mark("stage");

  BoxRun r;
  r.d = d; r.mu = i_mu.dev; r.U = i_U.dev; r.u_upper = u_upper; r.Sigma = i_S.dev; r.lb = i_lb.dev; r.ub = i_ub.dev;
  r.init = i_init.dev; r.init_ld = init_ld; r.n_chains = n_chains; r.n_per_chain = n_per_chain; r.burn_in = burn_in;
  r.seed = o.seed; r.chain_offset = o.chain_offset; r.prepass = o.prepass; r.max_restarts = o.max_restarts;
  r.exact_search = o.hmc_search == 1; r.no_window = o.hmc_search == 2; r.fork = o.fork; r.fork_burn_in = o.fork_burn_in;
  r.injected = o.injected ? i_inj.dev : nullptr; r.n_injected = o.n_injected;
  r.out = d_out; r.st = st;

  // HOST mode: samples every chain has already emitted are copied out while later rounds run
  // (one strided 2-D copy per round: chain-major columns, pitch = n_per_chain * d doubles).
  // A pageable result matrix (R, numpy) is filled through the pinned staging ring and helper
  // threads; a page-locked one (lqg_alloc_host) directly by the copy engine.  LQG_STAGED=0 disables.
  StreamGuard copy_guard;
  std::unique_ptr<StagedCopier> staged;
  if (mem == LQG_MEM_HOST) {
    LQG_CUDA_TRY(cudaStreamCreateWithFlags(&copy_guard.s, cudaStreamNonBlocking));
    static const char* env_staged = getenv("LQG_STAGED");
    if (!(env_staged && env_staged[0] == '0') && StagedCopier::is_pageable(out)) staged.reset(new StagedCopier(copy_guard.s));
    cudaStream_t copy_st = copy_guard.s;
    r.short_tail = true;
    r.on_round = [&, copy_st](int from, int upto, cudaEvent_t ready) -> int {
      LQG_CUDA_TRY(cudaStreamWaitEvent(copy_st, ready, 0));
      const size_t pitch = (size_t)n_per_chain * d * sizeof(double);
      const size_t width = (size_t)(upto - from) * d * sizeof(double);
      if (staged) LQG_CUDA_TRY(staged->copy2d(out + (size_t)from * d, pitch, d_out + (size_t)from * d, pitch, width, n_chains));
      else LQG_CUDA_TRY(cudaMemcpy2DAsync(out + (size_t)from * d, pitch, d_out + (size_t)from * d, pitch, width, n_chains,
                                          cudaMemcpyDeviceToHost, copy_st));
      return LQG_OK;
    };
  }
  BoxResult res;
  if ((rc = hmc_box_core(r, res))) return rc;
  if (copy_guard.s) {
    if (staged) LQG_CUDA_TRY(staged->finish());
    LQG_CUDA_TRY(cudaStreamSynchronize(copy_guard.s));
  }
  if (state_out) {
    // state_out holds positions in the ORIGINAL frame: mu + x_b, chain by chain
    std::vector<double> h_state((size_t)d * n_chains), h_mu;
    LQG_CUDA_TRY(cudaMemcpyAsync(h_state.data(), res.state.p, h_state.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    if ((rc = fetch_host(h_mu, mu, d, mem, st))) return rc;
    LQG_CUDA_TRY(cudaStreamSynchronize(st));
    for (int ch = 0; ch < n_chains; ++ch)
      for (int i = 0; i < d; ++i) h_state[(size_t)ch * d + i] += h_mu[i];
    if (mem == LQG_MEM_HOST) memcpy(state_out, h_state.data(), h_state.size() * sizeof(double));
    else LQG_CUDA_TRY(cudaMemcpyAsync(state_out, h_state.data(), h_state.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    LQG_CUDA_TRY(cudaStreamSynchronize(st));
  }
  ph.mark("copy out");
  rc = box_result_rc(res, "lqg_hmc_box");
  fill_stats(stats, res.stats, res.chain_ms, res.pre_ms, res.stats[10], res.stats[11]);
  fill_root_stats(stats, res.root_stats, res.root_chain_ms, res.root_pre_ms);
  fold_launches();
  return rc;
  LQG_API_END("lqg_hmc_box")
}

// ===========================================================================================
static int hmc_canonical_impl(int32_t d, int32_t c, const double* F, const double* g,
                              const double* initial, int64_t initial_ld, int32_t n_per_chain,
                              int32_t burn_in, const lqg_opts* opts_, double* out,
                              double* state_out, lqg_hmc_stats* stats, int32_t trace_cap,
                              double* trace_t, int32_t* trace_cn, int32_t* trace_n) {
  if (stats) memset(stats, 0, sizeof(*stats));
  if (!have_device()) return LQG_ERR_CUDA;
  const lqg_opts o = opts_ ? *opts_ : default_opts();
  const int n_chains = o.n_chains > 0 ? o.n_chains : 1;
  if (d <= 0 || d > 8192 || c < 0 || (c > 0 && (!F || !g)) || !initial || !out || n_per_chain <= 0 ||
      burn_in < 0 || (initial_ld != 0 && initial_ld < d)) {
    set_error("lqg_hmc_canonical: invalid argument (d=%d, c=%d, n_per_chain=%d)", d, c, n_per_chain);
    return LQG_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)o.stream;
  DevBuf::cur_stream = st;
  const int mem = o.mem;
  const size_t n_init = initial_ld == 0 ? (size_t)d : (size_t)initial_ld * (n_chains - 1) + d;
  const size_t n_F = (size_t)c * d;
  int rc;

  Input i_F, i_g, i_init, i_inj;
  // F is always copied: the bulk copy into shared memory needs a 16-byte multiple
  if ((rc = i_F.stage(F, n_F, mem, st, ((n_F + 1) & ~(size_t)1) + 2)) || (rc = i_g.stage(g, c, mem, st, c + 2)) ||
      (rc = i_init.stage(initial, n_init, mem, st)))
    return rc;
  if (o.injected && (rc = i_inj.stage(o.injected, (size_t)n_chains * o.n_injected, mem, st))) return rc;

  const size_t n_cols = (size_t)n_chains * n_per_chain;
  DevBuf b_out, b_state, b_f2, b_fn, b_misc, b_status, b_injpos;
  double* d_out = out;
  if (mem == LQG_MEM_HOST) { LQG_CUDA_TRY(b_out.alloc(n_cols * d * sizeof(double))); d_out = b_out.as<double>(); }
  LQG_CUDA_TRY(b_state.alloc((size_t)d * n_chains * sizeof(double)));
  LQG_CUDA_TRY(b_f2.alloc((c + 1) * sizeof(double)));
  LQG_CUDA_TRY(b_fn.alloc((c + 1) * sizeof(double)));
  LQG_CUDA_TRY(b_misc.alloc(16 * sizeof(unsigned long long)));
  LQG_CUDA_TRY(b_status.alloc((size_t)n_chains * sizeof(int)));
  LQG_CUDA_TRY(b_injpos.alloc((size_t)n_chains * sizeof(int64_t)));
  LQG_CUDA_TRY(cudaMemsetAsync(b_misc.p, 0, b_misc.bytes, st));
  LQG_CUDA_TRY(cudaMemsetAsync(b_status.p, 0, b_status.bytes, st));
  LQG_CUDA_TRY(cudaMemsetAsync(b_injpos.p, 0, b_injpos.bytes, st));
  if (c > 0) LQG_CUDA_TRY(launch_canon_rownorms(d, c, i_F.dev, b_f2.as<double>(), b_fn.as<double>(), st));
  // state = initial (canonical frame): reuse the box helper with mu = 0
  {
    DevBuf zero;
    LQG_CUDA_TRY(zero.alloc(d * sizeof(double)));
    LQG_CUDA_TRY(cudaMemsetAsync(zero.p, 0, d * sizeof(double), st));
    LQG_CUDA_TRY(launch_box_init_state(d, n_chains, i_init.dev, initial_ld, zero.as<double>(), b_state.as<double>(), st));
    LQG_CUDA_TRY(cudaStreamSynchronize(st));
  }

  CanonParams p;
  memset(&p, 0, sizeof(p));
  p.d = d; p.c = c; p.F = i_F.dev; p.g = i_g.dev; p.frow2 = b_f2.as<double>(); p.fnorm = b_fn.as<double>();
  p.state = b_state.as<double>(); p.n_chains = n_chains; p.s_begin = 0; p.s_end = burn_in + n_per_chain;
  p.burn_in = burn_in; p.n_per_chain = n_per_chain; p.out = d_out; p.seed = o.seed;
  p.chain_offset = o.chain_offset; p.injected = o.injected ? i_inj.dev : nullptr;
  p.n_injected = o.n_injected; p.inj_pos = b_injpos.as<int64_t>();
  p.max_restarts = o.max_restarts > 0 ? o.max_restarts : 100000;
  p.f_in_smem = (c > 0 && canon_f_fits_smem(d, c)) ? 1 : 0;
  p.stats = b_misc.as<unsigned long long>();
  p.next_chain = reinterpret_cast<int*>(b_misc.as<unsigned long long>() + 8);
  p.chain_status = b_status.as<int>();

  DevBuf b_tt, b_tc, b_tn;
  if (trace_cap > 0 && trace_t && trace_cn && trace_n) {
    LQG_CUDA_TRY(b_tt.alloc((size_t)n_chains * trace_cap * sizeof(double)));
    LQG_CUDA_TRY(b_tc.alloc((size_t)n_chains * trace_cap * sizeof(int)));
    LQG_CUDA_TRY(b_tn.alloc((size_t)n_chains * sizeof(int)));
    LQG_CUDA_TRY(cudaMemsetAsync(b_tt.p, 0, b_tt.bytes, st));
    LQG_CUDA_TRY(cudaMemsetAsync(b_tc.p, 0xff, b_tc.bytes, st));
    LQG_CUDA_TRY(cudaMemsetAsync(b_tn.p, 0, b_tn.bytes, st));
    p.trace_t = b_tt.as<double>(); p.trace_cn = b_tc.as<int>(); p.trace_cap = trace_cap; p.trace_n = b_tn.as<int>();
  }
  // large dense F (does not fit in shared memory) and enough chains: lock-step rounds whose dot
  // products are one tensor-core GEMM per round (lqg_hmc_canonical_batched.cu); lqg_opts.canon_mode overrides
  const bool batched = c > 0 && !p.trace_t &&
                       (o.canon_mode == 2 || (o.canon_mode == 0 && !p.f_in_smem && n_chains >= 64 && (size_t)c * d >= ((size_t)1 << 20)));
  Timer t_chain(st);
  DevBuf b_ft, b_ab, b_fab, b_dbl, b_int;
  if (batched) {
    const int n = n_chains;
    LQG_CUDA_TRY(b_ft.alloc(n_F * sizeof(double)));
    LQG_CUDA_TRY(b_ab.alloc((size_t)d * 2 * n * sizeof(double)));
    LQG_CUDA_TRY(b_fab.alloc((size_t)c * 2 * n * sizeof(double)));
    LQG_CUDA_TRY(b_dbl.alloc((size_t)2 * n * sizeof(double)));
    LQG_CUDA_TRY(b_int.alloc(((size_t)3 * n + 4) * sizeof(int)));
    LQG_CUDA_TRY(cudaMemsetAsync(b_ab.p, 0, b_ab.bytes, st));
    LQG_CUDA_TRY(cudaMemsetAsync(b_dbl.p, 0, b_dbl.bytes, st));
    LQG_CUDA_TRY(cudaMemsetAsync(b_int.p, 0, b_int.bytes, st));
    LQG_CUDA_TRY(launch_transpose(c, d, i_F.dev, c, b_ft.as<double>(), d, st));
    CanonBatch q;
    memset(&q, 0, sizeof(q));
    q.d = d; q.c = c; q.n = n; q.Ft = b_ft.as<double>(); q.g = i_g.dev; q.frow2 = p.frow2; q.fnorm = p.fnorm;
    q.AB = b_ab.as<double>(); q.Zlast = b_state.as<double>(); q.FAB = b_fab.as<double>();
    q.tt = b_dbl.as<double>(); q.E = q.tt + n;
    q.phase = b_int.as<int>(); q.s_done = q.phase + n; q.restart = q.s_done + n; q.n_unfinished = q.restart + n;
    q.pos = b_injpos.as<int64_t>(); q.status = b_status.as<int>();
    q.burn_in = burn_in; q.n_per_chain = n_per_chain; q.total = burn_in + n_per_chain; q.out = d_out;
    q.seed = o.seed; q.chain_offset = o.chain_offset; q.injected = p.injected; q.n_injected = p.n_injected;
    q.max_restarts = p.max_restarts; q.stats = p.stats;
    t_chain.start();
    for (long long round = 0;; ++round) {
      LQG_CUDA_TRY(cudaMemsetAsync(q.n_unfinished, 0, sizeof(int), st));
      LQG_CUDA_TRY(launch_canon_batch_round_begin(q, st));
      LQG_CUDA_TRY(launch_dgemm(c, 2 * n, d, i_F.dev, c, q.AB, d, b_fab.as<double>(), c, nullptr, 0, st));
      LQG_CUDA_TRY(launch_canon_batch_round_step(q, st));
      if ((round & 3) != 3) continue;                    // completion is polled every 4th round (idle rounds are no-ops)
      int unfinished = 0;
      LQG_CUDA_TRY(cudaMemcpyAsync(&unfinished, q.n_unfinished, sizeof(int), cudaMemcpyDeviceToHost, st));
      LQG_CUDA_TRY(cudaStreamSynchronize(st));
      if (unfinished == 0) break;
      if (round > 100000000ll) { set_error("lqg_hmc_canonical: round limit reached"); return LQG_ERR_INTERNAL; }
    }
    t_chain.stop();
  } else {
    t_chain.start();
    LQG_CUDA_TRY(launch_hmc_canonical(p, 0, st));
    t_chain.stop();
  }
  const double chain_ms = t_chain.ms();
  if (p.trace_t) {
    LQG_CUDA_TRY(cudaMemcpyAsync(trace_t, b_tt.p, (size_t)n_chains * trace_cap * sizeof(double), cudaMemcpyDeviceToHost, st));
    LQG_CUDA_TRY(cudaMemcpyAsync(trace_cn, b_tc.p, (size_t)n_chains * trace_cap * sizeof(int), cudaMemcpyDeviceToHost, st));
    LQG_CUDA_TRY(cudaMemcpyAsync(trace_n, b_tn.p, (size_t)n_chains * sizeof(int), cudaMemcpyDeviceToHost, st));
  }

  if (c > 0)
    LQG_CUDA_TRY(launch_canon_violations(d, c, n_cols, i_F.dev, i_g.dev, b_fn.as<double>(), d_out, p.stats + 6, st));
  unsigned long long h_stats[8];
  std::vector<int> h_status(n_chains);
  LQG_CUDA_TRY(cudaMemcpyAsync(h_stats, p.stats, sizeof(h_stats), cudaMemcpyDeviceToHost, st));
  LQG_CUDA_TRY(cudaMemcpyAsync(h_status.data(), b_status.p, n_chains * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (mem == LQG_MEM_HOST)
    LQG_CUDA_TRY(cudaMemcpyAsync(out, d_out, n_cols * d * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (state_out)
    LQG_CUDA_TRY(cudaMemcpyAsync(state_out, b_state.p, (size_t)d * n_chains * sizeof(double),
                                 mem == LQG_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, st));
  LQG_CUDA_TRY(cudaStreamSynchronize(st));
  rc = chain_status_rc(h_status, &h_stats[7]);
  fill_stats(stats, h_stats, chain_ms, 0.0);
  fold_launches();
  if (rc == LQG_ERR_RESTART_CAP) set_error("lqg_hmc_canonical: %llu chain(s) exceeded max_restarts=%d", h_stats[7], p.max_restarts);
  if (rc == LQG_ERR_DRAWS_EXHAUSTED) set_error("lqg_hmc_canonical: injected Gaussian stream too short for %llu chain(s)", h_stats[7]);
  return rc;
}

extern "C" int lqg_hmc_canonical(int32_t d, int32_t c, const double* F, const double* g,
                                 const double* initial, int64_t initial_ld, int32_t n_per_chain,
                                 int32_t burn_in, const lqg_opts* opts_, double* out,
                                 double* state_out, lqg_hmc_stats* stats) {
  LQG_API_BEGIN
  return hmc_canonical_impl(d, c, F, g, initial, initial_ld, n_per_chain, burn_in, opts_, out, state_out, stats,
                            0, nullptr, nullptr, nullptr);
  LQG_API_END("lqg_hmc_canonical")
}

extern "C" int lqg_hmc_canonical_trace(int32_t d, int32_t c, const double* F, const double* g,
                                       const double* initial, int64_t initial_ld, int32_t n_per_chain,
                                       const lqg_opts* opts_, double* out, int32_t trace_cap,
                                       double* trace_t, int32_t* trace_cn, int32_t* trace_n) {
  LQG_API_BEGIN
  if (opts_ && opts_->mem != LQG_MEM_HOST) { set_error("lqg_hmc_canonical_trace: host pointers only"); return LQG_ERR_INVALID; }
  if (trace_cap <= 0 || !trace_t || !trace_cn || !trace_n) { set_error("lqg_hmc_canonical_trace: invalid trace buffers"); return LQG_ERR_INVALID; }
  return hmc_canonical_impl(d, c, F, g, initial, initial_ld, n_per_chain, 0, opts_, out, nullptr, nullptr,
                            trace_cap, trace_t, trace_cn, trace_n);
  LQG_API_END("lqg_hmc_canonical_trace")
}

// ===========================================================================================
extern "C" int lqg_rtmg(int32_t n, int32_t seed, const double* initial, int32_t d, int32_t numlin,
                        const double* F, const double* g, int32_t numquad, const double* quadratics,
                        double* out) {
  LQG_API_BEGIN
  (void)quadratics;
  if (numquad != 0) {
    set_error("lqg_rtmg: quadratic constraints are not supported (lineqGPR always passes q=NULL)");
    return LQG_ERR_UNSUPPORTED;
  }
  if (n <= 0 || d <= 0 || !initial || !out) { set_error("lqg_rtmg: invalid argument"); return LQG_ERR_INVALID; }
  lqg_opts o = default_opts();
  o.seed = (uint64_t)(uint32_t)seed;
  o.n_chains = 1;
  std::vector<double> cols((size_t)n * d);
  int rc = lqg_hmc_canonical(d, numlin, F, g, initial, 0, n, 0, &o, cols.data(), nullptr, nullptr);
  if (rc != LQG_OK) return rc;
  for (int i = 0; i < n; ++i)                    // d x n (column = sample) -> n x d (row = sample)
    for (int j = 0; j < d; ++j) out[(size_t)i + (size_t)j * n] = cols[(size_t)i * d + j];
  return LQG_OK;
  LQG_API_END("lqg_rtmg")
}

// ===========================================================================================
extern "C" void* lqg_alloc_host(uint64_t bytes) {
  if (!have_device()) return nullptr;
  void* p = nullptr;
  cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 16, cudaHostAllocDefault);
  if (e != cudaSuccess) { cudaGetLastError(); set_error("lqg_alloc_host(%llu): %s", (unsigned long long)bytes, cudaGetErrorString(e)); return nullptr; }
  return p;
}
extern "C" void lqg_free_host(void* p) { if (p) cudaFreeHost(p); }

extern "C" const char* lqg_version(void) { return "lineqgpr_b200 0.1.0 (sm_100a)"; }
extern "C" const char* lqg_last_error(void) { return g_err; }
extern "C" int lqg_hmc_box_slots(int32_t d) {
  if (!have_device() || d <= 0 || d > 8192) return 0;
  return hmc_box_slots(d);
}
extern "C" int lqg_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}
// Parity tooling (see the header): the device hit-time functions over arrays, host pointers
extern "C" int lqg_hit_time_batch(int64_t n, const double* fa, const double* fb, const double* g, double klim,
                                  double* t_exact, int32_t* fast_class, double* fast_key, double* fast_sin, double* fast_cos) {
  LQG_API_BEGIN
  if (!have_device()) return LQG_ERR_CUDA;
  if (n <= 0 || !fa || !fb || !g || (!t_exact && !fast_class)) { set_error("lqg_hit_time_batch: invalid argument"); return LQG_ERR_INVALID; }
  cudaStream_t st = nullptr;
  DevBuf::cur_stream = st;
  const size_t nb = (size_t)n * sizeof(double);
  DevBuf b_fa, b_fb, b_g, b_t, b_c, b_k, b_s, b_co;
  LQG_CUDA_TRY(b_fa.alloc(nb)); LQG_CUDA_TRY(b_fb.alloc(nb)); LQG_CUDA_TRY(b_g.alloc(nb));
  LQG_CUDA_TRY(cudaMemcpyAsync(b_fa.p, fa, nb, cudaMemcpyHostToDevice, st));
  LQG_CUDA_TRY(cudaMemcpyAsync(b_fb.p, fb, nb, cudaMemcpyHostToDevice, st));
  LQG_CUDA_TRY(cudaMemcpyAsync(b_g.p, g, nb, cudaMemcpyHostToDevice, st));
  if (t_exact) LQG_CUDA_TRY(b_t.alloc(nb));
  if (fast_class) { LQG_CUDA_TRY(b_c.alloc((size_t)n * sizeof(int))); LQG_CUDA_TRY(b_k.alloc(nb)); LQG_CUDA_TRY(b_s.alloc(nb)); LQG_CUDA_TRY(b_co.alloc(nb)); }
  LQG_CUDA_TRY(launch_hit_time_batch((size_t)n, b_fa.as<double>(), b_fb.as<double>(), b_g.as<double>(), klim,
                                     t_exact ? b_t.as<double>() : nullptr, fast_class ? b_c.as<int>() : nullptr,
                                     b_k.as<double>(), b_s.as<double>(), b_co.as<double>(), st));
  if (t_exact) LQG_CUDA_TRY(cudaMemcpyAsync(t_exact, b_t.p, nb, cudaMemcpyDeviceToHost, st));
  if (fast_class) {
    LQG_CUDA_TRY(cudaMemcpyAsync(fast_class, b_c.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (fast_key) LQG_CUDA_TRY(cudaMemcpyAsync(fast_key, b_k.p, nb, cudaMemcpyDeviceToHost, st));
    if (fast_sin) LQG_CUDA_TRY(cudaMemcpyAsync(fast_sin, b_s.p, nb, cudaMemcpyDeviceToHost, st));
    if (fast_cos) LQG_CUDA_TRY(cudaMemcpyAsync(fast_cos, b_co.p, nb, cudaMemcpyDeviceToHost, st));
  }
  LQG_CUDA_TRY(cudaStreamSynchronize(st));
  fold_launches();
  return LQG_OK;
  LQG_API_END("lqg_hit_time_batch")
}

// diagnostic: per-phase SM-clock sums of the box kernel's hit search since the last call (zeros unless
// the library was built with -DLQG_BOX_TIMING; scripts/probe_box_phases.py).  Not part of the public header.
extern "C" int lqg_debug_box_phases(unsigned long long* out16) {
  if (!out16) return LQG_ERR_INVALID;
  LQG_CUDA_TRY(box_phase_cycles(out16));
  return LQG_OK;
}
extern "C" int lqg_guard_violations(void) { return g_guard_violations.load(); }
extern "C" int64_t lqg_launch_count(int32_t reset) {
  fold_launches();
  int64_t v = g_launches.load();
  if (reset) g_launches = 0;
  return v;
}
extern "C" int lqg_measure_fp64_peaks(double* dfma_tflops, double* dmma_tflops, const lqg_opts* opts) {
  LQG_API_BEGIN
  if (!have_device()) return LQG_ERR_CUDA;
  cudaStream_t st = opts ? (cudaStream_t)opts->stream : nullptr;
  LQG_CUDA_TRY(measure_fp64_peaks(dfma_tflops, dmma_tflops, st));
  fold_launches();
  return LQG_OK;
  LQG_API_END("lqg_measure_fp64_peaks")
}

// probe: DFMA warps and DMMA warps side by side (are the two FP64 paths separate units?)
extern "C" int lqg_measure_fp64_mixed(double* tflops, double* dfma_share, const lqg_opts* opts) {
  LQG_API_BEGIN
  if (!have_device()) return LQG_ERR_CUDA;
  cudaStream_t st = opts ? (cudaStream_t)opts->stream : nullptr;
  LQG_CUDA_TRY(measure_fp64_mixed(tflops, dfma_share, st));
  fold_launches();
  return LQG_OK;
  LQG_API_END("lqg_measure_fp64_mixed")
}

// ===========================================================================================
extern "C" int lqg_rsm(int32_t d, const double* map, const double* factor, const double* w,
                       const double* lb, const double* ub, int64_t nsim, int64_t max_proposals,
                       const lqg_opts* opts_, double* out, lqg_rsm_stats* stats) {
  LQG_API_BEGIN
  if (stats) memset(stats, 0, sizeof(*stats));
  if (!have_device()) return LQG_ERR_CUDA;
  const lqg_opts o = opts_ ? *opts_ : default_opts();
  if (d <= 0 || !map || !factor || !w || !lb || !ub || !out || nsim <= 0) {
    set_error("lqg_rsm: invalid argument (d=%d, nsim=%lld)", d, (long long)nsim);
    return LQG_ERR_INVALID;
  }
  // rsm_eval/rsm_emit keep one proposal vector per warp in shared memory: 8 warps x d doubles
  // must fit the 227 KB a CTA can opt into (rejection from the mode is hopeless long before that:
  // the acceptance rate falls geometrically with d)
  if ((size_t)d * 64 > (size_t)227 * 1024) {
    set_error("lqg_rsm: d=%d exceeds the supported maximum of %d (8 proposal vectors per CTA in shared memory)", d,
              (int)((size_t)227 * 1024 / 64));
    return LQG_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)o.stream;
  DevBuf::cur_stream = st;
  const int mem = o.mem;
  int rc;
  Input i_map, i_fac, i_w, i_lb, i_ub, i_z, i_u;
  if ((rc = i_map.stage(map, d, mem, st)) || (rc = i_fac.stage(factor, (size_t)d * d, mem, st)) ||
      (rc = i_w.stage(w, d, mem, st)) || (rc = i_lb.stage(lb, d, mem, st)) || (rc = i_ub.stage(ub, d, mem, st)))
    return rc;
  const bool inj = o.injected != nullptr;
  if (inj && (rc = i_z.stage(o.injected, (size_t)d * o.n_injected, mem, st))) return rc;
  if (o.injected_u && (rc = i_u.stage(o.injected_u, (size_t)o.n_injected_u, mem, st))) return rc;

  int64_t budget = max_proposals > 0 ? max_proposals : ((int64_t)1 << 40);
  if (inj && o.n_injected < budget) budget = o.n_injected;
  const int64_t kMaxRound = (int64_t)1 << 22;
  int64_t round = std::min<int64_t>(kMaxRound, std::max<int64_t>(4096, 4 * nsim));
  round = std::min(round, budget);
  if (round <= 0) { set_error("lqg_rsm: empty proposal budget"); return LQG_ERR_BUDGET; }

  DevBuf b_out, b_flags, b_t, b_cnt, b_ioff, b_aoff, b_ctr;
  double* d_out = out;
  if (mem == LQG_MEM_HOST) { LQG_CUDA_TRY(b_out.alloc((size_t)nsim * d * sizeof(double))); d_out = b_out.as<double>(); }
  const int nb_max = rsm_blocks(kMaxRound);
  LQG_CUDA_TRY(b_flags.alloc((size_t)kMaxRound));
  LQG_CUDA_TRY(b_t.alloc((size_t)kMaxRound * sizeof(double)));
  LQG_CUDA_TRY(b_cnt.alloc((size_t)nb_max * sizeof(int)));
  LQG_CUDA_TRY(b_ioff.alloc((size_t)(nb_max + 1) * sizeof(int64_t)));
  LQG_CUDA_TRY(b_aoff.alloc((size_t)(nb_max + 1) * sizeof(int64_t)));
  LQG_CUDA_TRY(b_ctr.alloc(2 * sizeof(unsigned long long)));
  LQG_CUDA_TRY(cudaMemsetAsync(b_ctr.p, 0, 2 * sizeof(unsigned long long), st));

  RsmParams q;
  memset(&q, 0, sizeof(q));
  q.d = d; q.map = i_map.dev; q.factor = i_fac.dev; q.w = i_w.dev; q.lb = i_lb.dev; q.ub = i_ub.dev;
  q.seed = o.seed; q.inj_z = inj ? i_z.dev : nullptr; q.inj_u = o.injected_u ? i_u.dev : nullptr;
  q.n_inj_u = o.n_injected_u; q.flags = b_flags.as<unsigned char>(); q.tval = b_t.as<double>();

  int64_t got = 0, first = 0, inbox = 0, rounds = 0;
  Timer tm(st);
  tm.start();
  while (got < nsim) {
    if (first >= budget) break;
    const int64_t P = std::min(round, budget - first);
    q.first_proposal = first; q.n_proposals = P; q.inbox_before = inbox;
    LQG_CUDA_TRY(launch_rsm_round(q, b_cnt.as<int>(), b_ioff.as<int64_t>(), b_aoff.as<int64_t>(), got, nsim,
                                  d_out, b_ctr.as<unsigned long long>(), st));
    const int nb = rsm_blocks(P);
    int64_t tot_in = 0, tot_acc = 0;
    LQG_CUDA_TRY(cudaMemcpyAsync(&tot_in, b_ioff.as<int64_t>() + nb, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    LQG_CUDA_TRY(cudaMemcpyAsync(&tot_acc, b_aoff.as<int64_t>() + nb, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    LQG_CUDA_TRY(cudaStreamSynchronize(st));
    got += tot_acc; inbox += tot_in; first += P; ++rounds;
    if (q.inj_u && inbox > q.n_inj_u && got < nsim) break;     // uniforms ran out
    const double rate = (double)std::max<int64_t>(got, 1) / (double)first;
    round = (int64_t)std::min<double>((double)kMaxRound, std::max<double>(4096.0, (double)(nsim - got) / rate * 1.3 + 1024.0));
  }
  tm.stop();
  const double ms = tm.ms();
  unsigned long long ctr[2] = {0, 0};
  LQG_CUDA_TRY(cudaMemcpyAsync(ctr, b_ctr.p, sizeof(ctr), cudaMemcpyDeviceToHost, st));
  const int64_t n_emit = std::min(got, nsim);
  if (mem == LQG_MEM_HOST && n_emit > 0)
    LQG_CUDA_TRY(cudaMemcpyAsync(out, d_out, (size_t)n_emit * d * sizeof(double), cudaMemcpyDeviceToHost, st));
  LQG_CUDA_TRY(cudaStreamSynchronize(st));
  if (stats) {
    stats->proposals = got >= nsim ? (int64_t)ctr[0] : first;
    stats->in_box = got >= nsim ? (int64_t)ctr[1] : inbox;
    stats->accepted = n_emit; stats->rounds = rounds; stats->kernel_ms = ms;
  }
  fold_launches();
  if (got < nsim) {
    if (inj || o.injected_u) {
      set_error("lqg_rsm: injected draws exhausted after %lld proposals (%lld of %lld accepted)",
                (long long)first, (long long)got, (long long)nsim);
      return LQG_ERR_DRAWS_EXHAUSTED;
    }
    set_error("lqg_rsm: proposal budget %lld exhausted (%lld of %lld accepted, %lld in box): "
              "the box has negligible mass under N(map, Sigma)", (long long)budget, (long long)got,
              (long long)nsim, (long long)inbox);
    return LQG_ERR_BUDGET;
  }
  return LQG_OK;
  LQG_API_END("lqg_rsm")
}

// ===========================================================================================
extern "C" int lqg_expt(int32_t d, const double* L0, const double* l, const double* u, const double* mu,
                        double psistar, int64_t nsim, int64_t max_proposals, const lqg_opts* opts_,
                        double* Z_out, lqg_expt_stats* stats) {
  LQG_API_BEGIN
  if (stats) memset(stats, 0, sizeof(*stats));
  if (!have_device()) return LQG_ERR_CUDA;
  const lqg_opts o = opts_ ? *opts_ : default_opts();
  if (d <= 0 || d > 8192 || !L0 || !l || !u || !mu || !Z_out || nsim <= 0) {
    set_error("lqg_expt: invalid argument (d=%d, nsim=%lld)", d, (long long)nsim);
    return LQG_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)o.stream;
  DevBuf::cur_stream = st;
  const int mem = o.mem;
  int rc;
  Input i_L, i_l, i_u, i_mu;
  if ((rc = i_L.stage(L0, (size_t)d * d, mem, st)) || (rc = i_l.stage(l, d, mem, st)) ||
      (rc = i_u.stage(u, d, mem, st)) || (rc = i_mu.stage(mu, d, mem, st)))
    return rc;
  const int64_t budget = max_proposals > 0 ? max_proposals : ((int64_t)1 << 40);
  // round size: Z scratch (d doubles per proposal) stays below ~1 GB
  const int64_t kMaxRound = std::max<int64_t>(4096, std::min<int64_t>((int64_t)1 << 20, ((int64_t)1 << 30) / ((int64_t)d * 8)));
  int64_t round = std::min<int64_t>(kMaxRound, std::max<int64_t>(4096, 2 * nsim));
  round = std::min(round, budget);
  DevBuf b_Lt, b_out, b_Z, b_flags, b_cnt, b_off, b_ctr;
  double* d_out = Z_out;
  if (mem == LQG_MEM_HOST) { LQG_CUDA_TRY(b_out.alloc((size_t)nsim * d * sizeof(double))); d_out = b_out.as<double>(); }
  const int nb_max = expt_blocks(kMaxRound);
  LQG_CUDA_TRY(b_Lt.alloc((size_t)d * d * sizeof(double)));
  LQG_CUDA_TRY(b_Z.alloc((size_t)kMaxRound * d * sizeof(double)));
  LQG_CUDA_TRY(b_flags.alloc((size_t)kMaxRound));
  LQG_CUDA_TRY(b_cnt.alloc((size_t)nb_max * sizeof(int)));
  LQG_CUDA_TRY(b_off.alloc((size_t)(nb_max + 1) * sizeof(int64_t)));
  LQG_CUDA_TRY(b_ctr.alloc(2 * sizeof(unsigned long long)));
  LQG_CUDA_TRY(cudaMemsetAsync(b_ctr.p, 0, 2 * sizeof(unsigned long long), st));
  LQG_CUDA_TRY(launch_expt_transpose(d, i_L.dev, b_Lt.as<double>(), st));
  ExptParams q;
  memset(&q, 0, sizeof(q));
  q.d = d; q.Lt = b_Lt.as<double>(); q.l = i_l.dev; q.u = i_u.dev; q.mu = i_mu.dev; q.psistar = psistar;
  q.seed = o.seed; q.Z = b_Z.as<double>(); q.flags = b_flags.as<unsigned char>();
  int64_t got = 0, first = 0, rounds = 0;
  Timer tm(st);
  tm.start();
  while (got < nsim && first < budget) {
    const int64_t P = std::min(round, budget - first);
    q.first_proposal = first; q.n_proposals = P;
    LQG_CUDA_TRY(launch_expt_round(q, b_cnt.as<int>(), b_off.as<int64_t>(), got, nsim, d_out, b_ctr.as<unsigned long long>(), st));
    int64_t tot = 0;
    LQG_CUDA_TRY(cudaMemcpyAsync(&tot, b_off.as<int64_t>() + expt_blocks(P), sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    LQG_CUDA_TRY(cudaStreamSynchronize(st));
    got += tot; first += P; ++rounds;
    const double rate = (double)std::max<int64_t>(got, 1) / (double)first;
    round = (int64_t)std::min<double>((double)kMaxRound, std::max<double>(4096.0, (double)(nsim - got) / rate * 1.3 + 1024.0));
  }
  tm.stop();
  const double ms = tm.ms();
  unsigned long long ctr[2] = {0, 0};
  LQG_CUDA_TRY(cudaMemcpyAsync(ctr, b_ctr.p, sizeof(ctr), cudaMemcpyDeviceToHost, st));
  const int64_t n_emit = std::min(got, nsim);
  if (mem == LQG_MEM_HOST && n_emit > 0)
    LQG_CUDA_TRY(cudaMemcpyAsync(Z_out, d_out, (size_t)n_emit * d * sizeof(double), cudaMemcpyDeviceToHost, st));
  LQG_CUDA_TRY(cudaStreamSynchronize(st));
  if (stats) {
    stats->proposals = got >= nsim ? (int64_t)ctr[0] : first;
    stats->accepted = n_emit; stats->rounds = rounds; stats->kernel_ms = ms;
  }
  fold_launches();
  if (got < nsim) {
    set_error("lqg_expt: proposal budget %lld exhausted (%lld of %lld accepted)", (long long)budget,
              (long long)got, (long long)nsim);
    return LQG_ERR_BUDGET;
  }
  return LQG_OK;
  LQG_API_END("lqg_expt")
}

// ===========================================================================================
extern "C" int lqg_dgemm(int32_t m, int32_t n, int32_t k, const double* A, int64_t lda,
                         const double* B, int64_t ldb, double* C, int64_t ldc, double* row_sum,
                         const lqg_opts* opts_, double* kernel_ms) {
  LQG_API_BEGIN
  if (!have_device()) return LQG_ERR_CUDA;
  const lqg_opts o = opts_ ? *opts_ : default_opts();
  if (m <= 0 || n <= 0 || k <= 0 || !A || !B || (!C && !row_sum) || lda < m || ldb < k || (C && ldc < m)) {
    set_error("lqg_dgemm: invalid argument (m=%d n=%d k=%d lda=%lld ldb=%lld ldc=%lld)", m, n, k,
              (long long)lda, (long long)ldb, (long long)ldc);
    return LQG_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)o.stream;
  DevBuf::cur_stream = st;
  const int mem = o.mem;
  int rc;
  Input i_A, i_B;
  if ((rc = i_A.stage(A, (size_t)lda * (k - 1) + m, mem, st)) || (rc = i_B.stage(B, (size_t)ldb * (n - 1) + k, mem, st)))
    return rc;
  DevBuf b_C, b_rs;
  double* d_C = C; double* d_rs = row_sum;
  if (mem == LQG_MEM_HOST) {
    if (C) { LQG_CUDA_TRY(b_C.alloc(((size_t)ldc * (n - 1) + m) * sizeof(double))); d_C = b_C.as<double>(); }
    if (row_sum) { LQG_CUDA_TRY(b_rs.alloc(m * sizeof(double))); d_rs = b_rs.as<double>(); }
  }
  if (d_rs) LQG_CUDA_TRY(cudaMemsetAsync(d_rs, 0, m * sizeof(double), st));
  Timer tm(st);
  tm.start();
  LQG_CUDA_TRY(launch_dgemm(m, n, k, i_A.dev, lda, i_B.dev, ldb, d_C, ldc, d_rs, 0, st));
  tm.stop();
  if (kernel_ms) *kernel_ms = tm.ms();
  if (mem == LQG_MEM_HOST) {
    if (C) {
      if (ldc == m) LQG_CUDA_TRY(d2h_result(C, (size_t)m * n * sizeof(double), d_C, (size_t)m * n * sizeof(double), (size_t)m * n * sizeof(double), 1, st));
      else LQG_CUDA_TRY(d2h_result(C, ldc * sizeof(double), d_C, ldc * sizeof(double), m * sizeof(double), n, st));
    }
    if (row_sum) LQG_CUDA_TRY(cudaMemcpyAsync(row_sum, d_rs, m * sizeof(double), cudaMemcpyDeviceToHost, st));
  }
  LQG_CUDA_TRY(cudaStreamSynchronize(st));
  fold_launches();
  return LQG_OK;
  LQG_API_END("lqg_dgemm")
}

// ===========================================================================================
// Set-up stage on the device (lqg_linalg.cu)
constexpr int kInfoClean = 0x7f7f7f7f;

// copies a result back to the caller (host) or leaves it where it was computed (device)
static int deliver(double* dst, const double* src_dev, size_t n, int mem, cudaStream_t st) {
  if (!dst || dst == src_dev) return LQG_OK;
  LQG_CUDA_TRY(cudaMemcpyAsync(dst, src_dev, n * sizeof(double),
                               mem == LQG_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
  return LQG_OK;
}

// L (device, d x d) <- lower Cholesky factor of sym(A) (reversed index order when rev).
// Synchronises the stream.  *bad = 0 or first bad pivot.
static int chol_device(int d, const double* A_dev, int rev, double* L_dev, int* bad, cudaStream_t st) {
  DevBuf b_w, b_info;
  LQG_CUDA_TRY(b_w.alloc((size_t)d * d * sizeof(double)));
  LQG_CUDA_TRY(b_info.alloc(sizeof(int)));
  LQG_CUDA_TRY(cudaMemsetAsync(b_info.p, 0x7f, sizeof(int), st));
  LQG_CUDA_TRY(launch_sym_reverse(d, A_dev, d, b_w.as<double>(), d, rev, st));
  LQG_CUDA_TRY(launch_chol(d, b_w.as<double>(), d, L_dev, d, b_info.as<int>(), st));
  int info = 0;
  LQG_CUDA_TRY(cudaMemcpyAsync(&info, b_info.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  LQG_CUDA_TRY(cudaStreamSynchronize(st));
  *bad = info == kInfoClean ? 0 : info;
  return LQG_OK;
}

// U_dev (d x d) <- the upper factor with U U' = sym(Sigma):  U = J chol(J Sigma J) J
static int ul_factor_device(int d, const double* S_dev, double* U_dev, cudaStream_t st) {
  DevBuf b_l;
  LQG_CUDA_TRY(b_l.alloc((size_t)d * d * sizeof(double)));
  int bad = 0, rc;
  if ((rc = chol_device(d, S_dev, 1, b_l.as<double>(), &bad, st))) return rc;
  if (bad) {
    set_error("Error: M must be positive definite. (pivot %d of the reversed factorisation of Sigma is not positive)", bad);
    return LQG_ERR_NOT_PD;                                          // tmg:R/tmg.R:17-20
  }
  LQG_CUDA_TRY(launch_reverse(d, b_l.as<double>(), U_dev, st));
  return LQG_OK;
}

extern "C" int lqg_whiten_box(int32_t d, const double* Sigma, const lqg_opts* opts_, double* U_out,
                              double* kernel_ms) {
  LQG_API_BEGIN
  if (!have_device()) return LQG_ERR_CUDA;
  const lqg_opts o = opts_ ? *opts_ : default_opts();
  if (d <= 0 || d > 32768 || !Sigma || !U_out) { set_error("lqg_whiten_box: invalid argument (d=%d)", d); return LQG_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)o.stream;
  DevBuf::cur_stream = st;
  const size_t dd = (size_t)d * d;
  int rc;
  Input i_S;
  if ((rc = i_S.stage(Sigma, dd, o.mem, st))) return rc;
  DevBuf b_U;
  double* d_U = U_out;
  if (o.mem == LQG_MEM_HOST) { LQG_CUDA_TRY(b_U.alloc(dd * sizeof(double))); d_U = b_U.as<double>(); }
  Timer tm(st);
  tm.start();
  if ((rc = ul_factor_device(d, i_S.dev, d_U, st))) return rc;
  tm.stop();
  if (kernel_ms) *kernel_ms = tm.ms();
  if ((rc = deliver(U_out, d_U, dd, o.mem, st))) return rc;
  LQG_CUDA_TRY(cudaStreamSynchronize(st));
  fold_launches();
  return LQG_OK;
  LQG_API_END("lqg_whiten_box")
}

extern "C" int lqg_chol(int32_t d, const double* A, const lqg_opts* opts_, double* L_out, int32_t* bad_pivot) {
  LQG_API_BEGIN
  if (!have_device()) return LQG_ERR_CUDA;
  const lqg_opts o = opts_ ? *opts_ : default_opts();
  if (d <= 0 || d > 32768 || !A) { set_error("lqg_chol: invalid argument (d=%d)", d); return LQG_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)o.stream;
  DevBuf::cur_stream = st;
  const size_t dd = (size_t)d * d;
  int rc;
  Input i_A;
  if ((rc = i_A.stage(A, dd, o.mem, st))) return rc;
  DevBuf b_l;
  LQG_CUDA_TRY(b_l.alloc(dd * sizeof(double)));
  int bad = 0;
  if ((rc = chol_device(d, i_A.dev, 0, b_l.as<double>(), &bad, st))) return rc;
  if (bad_pivot) *bad_pivot = bad;
  fold_launches();
  if (bad) { set_error("lqg_chol: the matrix is not positive definite (pivot %d)", bad); return LQG_ERR_NOT_PD; }
  if ((rc = deliver(L_out, b_l.as<double>(), dd, o.mem, st))) return rc;
  LQG_CUDA_TRY(cudaStreamSynchronize(st));
  return LQG_OK;
  LQG_API_END("lqg_chol")
}

extern "C" int lqg_solve_spd(int32_t d, int32_t nrhs, const double* A, const double* B, const lqg_opts* opts_,
                             double* X_out) {
  LQG_API_BEGIN
  if (!have_device()) return LQG_ERR_CUDA;
  const lqg_opts o = opts_ ? *opts_ : default_opts();
  if (d <= 0 || d > 32768 || nrhs <= 0 || !A || !B || !X_out) {
    set_error("lqg_solve_spd: invalid argument (d=%d nrhs=%d)", d, nrhs);
    return LQG_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)o.stream;
  DevBuf::cur_stream = st;
  const size_t dd = (size_t)d * d, dn = (size_t)d * nrhs;
  int rc;
  Input i_A, i_B;
  if ((rc = i_A.stage(A, dd, o.mem, st)) || (rc = i_B.stage(B, dn, o.mem, st))) return rc;
  DevBuf b_l, b_x;
  LQG_CUDA_TRY(b_l.alloc(dd * sizeof(double)));
  LQG_CUDA_TRY(b_x.alloc(dn * sizeof(double)));
  int bad = 0;
  if ((rc = chol_device(d, i_A.dev, 0, b_l.as<double>(), &bad, st))) return rc;
  if (bad) { set_error("lqg_solve_spd: the matrix is not positive definite (pivot %d)", bad); fold_launches(); return LQG_ERR_NOT_PD; }
  LQG_CUDA_TRY(cudaMemcpyAsync(b_x.p, i_B.dev, dn * sizeof(double), cudaMemcpyDeviceToDevice, st));
  LQG_CUDA_TRY(launch_chol_solve(d, b_l.as<double>(), d, b_x.as<double>(), d, nrhs, st));
  if ((rc = deliver(X_out, b_x.as<double>(), dn, o.mem, st))) return rc;
  LQG_CUDA_TRY(cudaStreamSynchronize(st));
  fold_launches();
  return LQG_OK;
  LQG_API_END("lqg_solve_spd")
}

extern "C" int lqg_eta_params(int32_t d, int32_t m, const double* Lambda, const double* mu, const double* xi_map,
                              const double* Sigma, double nugget, const lqg_opts* opts_,
                              double* eta_mu, double* eta_map, double* Sigma_eta) {
  LQG_API_BEGIN
  if (!have_device()) return LQG_ERR_CUDA;
  const lqg_opts o = opts_ ? *opts_ : default_opts();
  if (d <= 0 || m <= 0 || d > 32768 || m > 32768 || !Lambda || !Sigma || !Sigma_eta || (eta_mu && !mu) || (eta_map && !xi_map)) {
    set_error("lqg_eta_params: invalid argument (d=%d m=%d)", d, m);
    return LQG_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)o.stream;
  DevBuf::cur_stream = st;
  const int mem = o.mem;
  const size_t dd = (size_t)d * d, dm = (size_t)d * m, mm = (size_t)m * m;
  int rc;
  Input i_L, i_S, i_mu, i_map;
  if ((rc = i_L.stage(Lambda, dm, mem, st)) || (rc = i_S.stage(Sigma, mm, mem, st))) return rc;
  if (eta_mu && (rc = i_mu.stage(mu, m, mem, st))) return rc;
  if (eta_map && (rc = i_map.stage(xi_map, m, mem, st))) return rc;
  DevBuf b_t, b_se, b_v;
  LQG_CUDA_TRY(b_t.alloc(dm * sizeof(double)));
  LQG_CUDA_TRY(b_v.alloc(2 * (size_t)d * sizeof(double)));
  double* d_se = Sigma_eta;
  if (mem == LQG_MEM_HOST) { LQG_CUDA_TRY(b_se.alloc(dd * sizeof(double))); d_se = b_se.as<double>(); }
  // T = Sigma Lambda' (m x d);  Sigma_eta = Lambda T + nugget I      R/lineqGP.R:613-614
  LQG_CUDA_TRY(launch_gemm(false, true, m, d, m, 1.0, i_S.dev, m, i_L.dev, d, 0.0, b_t.as<double>(), m, 0, st));
  LQG_CUDA_TRY(launch_gemm(false, false, d, d, m, 1.0, i_L.dev, d, b_t.as<double>(), m, 0.0, d_se, d, 0, st));
  if (nugget != 0.0) LQG_CUDA_TRY(launch_add_diag(d, d_se, d, nugget, st));
  double* v = b_v.as<double>();
  if (eta_mu) {                                                      // :611
    LQG_CUDA_TRY(launch_gemm(false, false, d, 1, m, 1.0, i_L.dev, d, i_mu.dev, m, 0.0, v, d, 0, st));
    if ((rc = deliver(eta_mu, v, d, mem, st))) return rc;
  }
  if (eta_map) {                                                     // :612
    LQG_CUDA_TRY(launch_gemm(false, false, d, 1, m, 1.0, i_L.dev, d, i_map.dev, m, 0.0, v + d, d, 0, st));
    if ((rc = deliver(eta_map, v + d, d, mem, st))) return rc;
  }
  if ((rc = deliver(Sigma_eta, d_se, dd, mem, st))) return rc;
  LQG_CUDA_TRY(cudaStreamSynchronize(st));
  fold_launches();
  return LQG_OK;
  LQG_API_END("lqg_eta_params")
}

// X (m x d, device) <- Lambda^+ for a device-resident Lambda (d x m, d >= m).  Synchronises the stream.
static int lambda_pinv_device(int d, int m, const double* L_dev, double* X, cudaStream_t st) {
  const size_t dd = (size_t)d * d, dm = (size_t)d * m, mm = (size_t)m * m;
  int rc;
  DevBuf b_g, b_l, b_r, b_y;
  LQG_CUDA_TRY(b_g.alloc(mm * sizeof(double)));
  LQG_CUDA_TRY(b_l.alloc(mm * sizeof(double)));
  LQG_CUDA_TRY(b_r.alloc(dd * sizeof(double)));
  LQG_CUDA_TRY(b_y.alloc(dm * sizeof(double)));
  double *G = b_g.as<double>(), *Lc = b_l.as<double>(), *R = b_r.as<double>(), *Y = b_y.as<double>();
  // G = Lambda' Lambda;  G = Lc Lc'
  LQG_CUDA_TRY(launch_gemm(true, false, m, m, d, 1.0, L_dev, d, L_dev, d, 0.0, G, m, 0, st));
  int bad = 0;
  if ((rc = chol_device(m, G, 0, Lc, &bad, st))) return rc;
  if (bad) {
    set_error("lqg_lambda_pinv: Lambda is rank deficient (pivot %d of Lambda'Lambda is not positive)", bad);
    return LQG_ERR_NOT_PD;
  }
  // X0 = G^-1 Lambda'
  LQG_CUDA_TRY(launch_transpose(d, m, L_dev, d, X, m, st));
  LQG_CUDA_TRY(launch_chol_solve(m, Lc, m, X, m, d, st));
  // one refinement step: X += G^-1 Lambda' (I - Lambda X0)
  LQG_CUDA_TRY(launch_gemm(false, false, d, d, m, -1.0, L_dev, d, X, m, 0.0, R, d, 0, st));
  LQG_CUDA_TRY(launch_add_diag(d, R, d, 1.0, st));
  LQG_CUDA_TRY(launch_gemm(true, false, m, d, d, 1.0, L_dev, d, R, d, 0.0, Y, m, 0, st));
  LQG_CUDA_TRY(launch_chol_solve(m, Lc, m, Y, m, d, st));
  LQG_CUDA_TRY(launch_axpy(dm, 1.0, Y, X, st));
  LQG_CUDA_TRY(cudaStreamSynchronize(st));                   // the scratch above dies with this frame
  return LQG_OK;
}

extern "C" int lqg_lambda_pinv(int32_t d, int32_t m, const double* Lambda, const lqg_opts* opts_, double* Lp_out) {
  LQG_API_BEGIN
  if (!have_device()) return LQG_ERR_CUDA;
  const lqg_opts o = opts_ ? *opts_ : default_opts();
  if (d <= 0 || m <= 0 || d < m || d > 32768 || !Lambda || !Lp_out) {
    set_error("lqg_lambda_pinv: invalid argument (d=%d m=%d; needs d >= m)", d, m);
    return LQG_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)o.stream;
  DevBuf::cur_stream = st;
  const int mem = o.mem;
  const size_t dm = (size_t)d * m;
  int rc;
  Input i_L;
  if ((rc = i_L.stage(Lambda, dm, mem, st))) return rc;
  DevBuf b_x;
  LQG_CUDA_TRY(b_x.alloc(dm * sizeof(double)));
  rc = lambda_pinv_device(d, m, i_L.dev, b_x.as<double>(), st);
  fold_launches();
  if (rc) return rc;
  if ((rc = deliver(Lp_out, b_x.as<double>(), dm, mem, st))) return rc;
  LQG_CUDA_TRY(cudaStreamSynchronize(st));
  return LQG_OK;
  LQG_API_END("lqg_lambda_pinv")
}

extern "C" int lqg_expt_setup(int32_t d, const double* Sigma, const double* l, const double* u, const lqg_opts* opts_,
                              double* L0, double* l_out, double* u_out, double* mu_out, double* x_out, double* psistar,
                              int32_t* perm, double* Lfull, int32_t* iters) {
  LQG_API_BEGIN
  if (!have_device()) return LQG_ERR_CUDA;
  const lqg_opts o = opts_ ? *opts_ : default_opts();
  if (d <= 0 || d > 2048 || !Sigma || !l || !u || !L0 || !l_out || !u_out || !mu_out || !psistar || !perm) {
    set_error("lqg_expt_setup: invalid argument (d=%d; needs 1 <= d <= 2048)", d);
    return LQG_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)o.stream;
  DevBuf::cur_stream = st;
  const int mem = o.mem;
  const size_t dd = (size_t)d * d;
  int rc;
  std::vector<double> h_l, h_u;
  if ((rc = fetch_host(h_l, l, d, mem, st)) || (rc = fetch_host(h_u, u, d, mem, st))) return rc;
  for (int i = 0; i < d; ++i)
    if (!(h_l[i] < h_u[i])) { set_error("l, u, and Sig have to match in dimension with u>l"); return LQG_ERR_INVALID; }
  DevBuf b_S, b_l, b_u, b_Lf, b_L0, b_perm, b_y, b_work, b_info;
  LQG_CUDA_TRY(b_S.alloc(dd * sizeof(double)));
  LQG_CUDA_TRY(b_l.alloc(d * sizeof(double)));
  LQG_CUDA_TRY(b_u.alloc(d * sizeof(double)));
  LQG_CUDA_TRY(b_Lf.alloc(dd * sizeof(double)));
  LQG_CUDA_TRY(b_L0.alloc(dd * sizeof(double)));
  LQG_CUDA_TRY(b_perm.alloc(d * sizeof(int)));
  LQG_CUDA_TRY(b_y.alloc((size_t)2 * d * sizeof(double)));
  LQG_CUDA_TRY(b_work.alloc((3 * dd + 32 * (size_t)d + 64) * sizeof(double)));
  LQG_CUDA_TRY(b_info.alloc(2 * sizeof(int)));
  const cudaMemcpyKind in_kind = mem == LQG_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  LQG_CUDA_TRY(cudaMemcpyAsync(b_S.p, Sigma, dd * sizeof(double), in_kind, st));
  LQG_CUDA_TRY(cudaMemcpyAsync(b_l.p, l, d * sizeof(double), in_kind, st));
  LQG_CUDA_TRY(cudaMemcpyAsync(b_u.p, u, d * sizeof(double), in_kind, st));
  LQG_CUDA_TRY(cudaMemsetAsync(b_y.p, 0, b_y.bytes, st));
  double ps = 0.0; int its = 0;
  const int r = expt_setup_device(d, b_S.as<double>(), b_l.as<double>(), b_u.as<double>(), b_Lf.as<double>(), b_L0.as<double>(),
                                  b_perm.as<int>(), b_y.as<double>(), b_work.as<double>(), b_info.as<int>(), &ps, &its, st);
  fold_launches();
  if (r < 0) { set_error("lqg_expt_setup: %s", cudaGetErrorString((cudaError_t)(-r))); return LQG_ERR_CUDA; }
  if (r == 1) { set_error("Sigma is not positive semi-definite"); return LQG_ERR_NOT_PD; }
  if (r == 2 || r == 3) { set_error("Covariance matrix is ill-conditioned and method failed"); return LQG_ERR_INVALID; }
  // tilt: y = [x (d-1) | mu (d-1)], both padded with a final zero
  std::vector<double> h_y((size_t)2 * d, 0.0), h_x(d, 0.0), h_mu(d, 0.0);
  if (d > 1) LQG_CUDA_TRY(cudaMemcpyAsync(h_y.data(), b_y.p, (size_t)2 * (d - 1) * sizeof(double), cudaMemcpyDeviceToHost, st));
  LQG_CUDA_TRY(cudaStreamSynchronize(st));
  for (int i = 0; i + 1 < d; ++i) { h_x[i] = h_y[i]; h_mu[i] = h_y[d - 1 + i]; }
  std::vector<int> h_perm(d);
  LQG_CUDA_TRY(cudaMemcpyAsync(h_perm.data(), b_perm.p, d * sizeof(int), cudaMemcpyDeviceToHost, st));
  LQG_CUDA_TRY(cudaStreamSynchronize(st));
  const cudaMemcpyKind hk = mem == LQG_MEM_DEVICE ? cudaMemcpyHostToDevice : cudaMemcpyHostToHost;
  LQG_CUDA_TRY(cudaMemcpyAsync(mu_out, h_mu.data(), d * sizeof(double), hk, st));
  if (x_out) LQG_CUDA_TRY(cudaMemcpyAsync(x_out, h_x.data(), d * sizeof(double), hk, st));
  LQG_CUDA_TRY(cudaMemcpyAsync(perm, h_perm.data(), d * sizeof(int), hk, st));
  if ((rc = deliver(L0, b_L0.as<double>(), dd, mem, st)) || (rc = deliver(l_out, b_l.as<double>(), d, mem, st)) ||
      (rc = deliver(u_out, b_u.as<double>(), d, mem, st)) || (Lfull && (rc = deliver(Lfull, b_Lf.as<double>(), dd, mem, st))))
    return rc;
  LQG_CUDA_TRY(cudaStreamSynchronize(st));
  *psistar = ps;
  if (iters) *iters = its;
  return LQG_OK;
  LQG_API_END("lqg_expt_setup")
}

extern "C" int lqg_basis(int32_t n_t, int32_t D, const int32_t* m_k, const double* x, int64_t ldx, const double* u,
                         const lqg_opts* opts_, double* Phi, int64_t ld) {
  LQG_API_BEGIN
  if (!have_device()) return LQG_ERR_CUDA;
  const lqg_opts o = opts_ ? *opts_ : default_opts();
  if (n_t <= 0 || D <= 0 || !m_k || !x || !u || !Phi || ldx < n_t || ld < n_t) {
    set_error("lqg_basis: invalid argument (n_t=%d D=%d ldx=%lld ld=%lld)", n_t, D, (long long)ldx, (long long)ld);
    return LQG_ERR_INVALID;
  }
  size_t mt = 0;
  for (int k = 0; k < D; ++k) {
    if (m_k[k] < 2) { set_error("lqg_basis: block %d has %d knots (needs >= 2)", k, m_k[k]); return LQG_ERR_INVALID; }
    mt += (size_t)m_k[k];
  }
  cudaStream_t st = (cudaStream_t)o.stream;
  DevBuf::cur_stream = st;
  const int mem = o.mem;
  int rc;
  Input i_x, i_u;
  if ((rc = i_x.stage(x, (size_t)ldx * (D - 1) + n_t, mem, st)) || (rc = i_u.stage(u, mt, mem, st))) return rc;
  DevBuf b_phi, b_flag;
  double* d_phi = Phi;
  const size_t n_phi = (size_t)ld * (mt - 1) + n_t;
  if (mem == LQG_MEM_HOST) { LQG_CUDA_TRY(b_phi.alloc(n_phi * sizeof(double))); d_phi = b_phi.as<double>(); }
  LQG_CUDA_TRY(b_flag.alloc(sizeof(int)));
  LQG_CUDA_TRY(cudaMemsetAsync(b_flag.p, 0, sizeof(int), st));
  size_t off = 0;
  for (int k = 0; k < D; ++k) {
    LQG_CUDA_TRY(launch_basis_hat(n_t, m_k[k], i_x.dev + (size_t)k * ldx, i_u.dev + off, d_phi + off * ld, ld,
                                  b_flag.as<int>(), st));
    off += (size_t)m_k[k];
  }
  int flag = 0;
  LQG_CUDA_TRY(cudaMemcpyAsync(&flag, b_flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  if (mem == LQG_MEM_HOST) {
    if (ld == n_t) LQG_CUDA_TRY(cudaMemcpyAsync(Phi, d_phi, n_phi * sizeof(double), cudaMemcpyDeviceToHost, st));
    else LQG_CUDA_TRY(cudaMemcpy2DAsync(Phi, ld * sizeof(double), d_phi, ld * sizeof(double), n_t * sizeof(double), mt, cudaMemcpyDeviceToHost, st));
  }
  LQG_CUDA_TRY(cudaStreamSynchronize(st));
  fold_launches();
  if (flag) {
    set_error("basisCompute.lineqGP: x outside the knot range [u_1, u_m]");
    return LQG_ERR_INVALID;
  }
  return LQG_OK;
  LQG_API_END("lqg_basis")
}

extern "C" int lqg_solve_qp(int32_t n, int32_t mc, const double* D, const double* dvec, const double* A,
                            const double* b, double tol, int32_t max_iter, const lqg_opts* opts_,
                            double* x_out, int32_t* iters) {
  LQG_API_BEGIN
  if (!have_device()) return LQG_ERR_CUDA;
  const lqg_opts o = opts_ ? *opts_ : default_opts();
  if (n <= 0 || n > 32768 || mc < 0 || !D || !dvec || !x_out || (mc > 0 && (!A || !b))) {
    set_error("lqg_solve_qp: invalid argument (n=%d mc=%d)", n, mc);
    return LQG_ERR_INVALID;
  }
  const bool polish = tol >= 0.0;           // tol < 0: |tol| is the tolerance and the active-set polish is skipped
  if (tol < 0.0) tol = -tol;
  if (tol == 0.0) tol = 1e-10;
  if (max_iter <= 0) max_iter = 200;
  cudaStream_t st = (cudaStream_t)o.stream;
  DevBuf::cur_stream = st;
  const int mem = o.mem;
  const size_t nn = (size_t)n * n, cn = (size_t)mc * n;
  int rc;
  std::vector<double> h_dvec, h_b;
  if ((rc = fetch_host(h_dvec, dvec, n, mem, st))) return rc;
  if (mc > 0 && (rc = fetch_host(h_b, b, mc, mem, st))) return rc;
  double scale_d = 1.0, scale_p = 1.0;
  for (double v : h_dvec) scale_d = std::max(scale_d, std::fabs(v));
  for (double v : h_b) {
    if (!std::isfinite(v)) { set_error("lqg_solve_qp: b must be finite (drop rows with infinite bounds)"); return LQG_ERR_INVALID; }
    scale_p = std::max(scale_p, std::fabs(v));
  }
  Input i_D, i_dv, i_A, i_b;
  if ((rc = i_D.stage(D, nn, mem, st)) || (rc = i_dv.stage(dvec, n, mem, st))) return rc;
  if (mc > 0 && ((rc = i_A.stage(A, cn, mem, st)) || (rc = i_b.stage(b, mc, mem, st)))) return rc;

  DevBuf b_L, b_H, b_AW, b_vn, b_vc, b_scal;
  LQG_CUDA_TRY(b_L.alloc(nn * sizeof(double)));
  LQG_CUDA_TRY(b_vn.alloc((size_t)3 * n * sizeof(double)));
  double* x = b_vn.as<double>(); double* rd = x + n; double* rhs = rd + n;
  double* Lc = b_L.as<double>();
  int bad = 0;
  // x = solve(D, dvec)
  if ((rc = chol_device(n, i_D.dev, 0, Lc, &bad, st))) return rc;
  if (bad) { set_error("lqg_solve_qp: Dmat is not positive definite (pivot %d)", bad); fold_launches(); return LQG_ERR_NOT_PD; }
  LQG_CUDA_TRY(cudaMemcpyAsync(x, i_dv.dev, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
  LQG_CUDA_TRY(launch_chol_solve(n, Lc, n, x, n, 1, st));
  int it = 0;
  bool converged = mc == 0;
  const double* x_result = x;
  DevBuf b_best, b_x0;
  if (mc > 0) {
    LQG_CUDA_TRY(b_best.alloc(((size_t)n + 2 * (size_t)mc) * sizeof(double)));   // best x, s, lam
    LQG_CUDA_TRY(b_x0.alloc(n * sizeof(double)));
    LQG_CUDA_TRY(cudaMemcpyAsync(b_x0.p, x, n * sizeof(double), cudaMemcpyDeviceToDevice, st));   // x0 = solve(D, dvec)
    double best_phi = INFINITY;
    int stall = 0;
    LQG_CUDA_TRY(b_H.alloc(nn * sizeof(double)));
    LQG_CUDA_TRY(b_AW.alloc(cn * sizeof(double)));
    LQG_CUDA_TRY(b_vc.alloc((size_t)8 * mc * sizeof(double)));
    LQG_CUDA_TRY(b_scal.alloc(8 * sizeof(double)));
    double* H = b_H.as<double>(); double* AW = b_AW.as<double>(); double* scal = b_scal.as<double>();
    double* s = b_vc.as<double>(); double* lam = s + mc; double* ds = lam + mc; double* dl = ds + mc;
    double* rp = dl + mc; double* t = rp + mc; double* rcv = t + mc; double* Ax = rcv + mc;
    const double* Ad = i_A.dev; const double* bd = i_b.dev;
    double h[8];
    auto read_scal = [&]() -> int {
      LQG_CUDA_TRY(cudaMemcpyAsync(h, scal, 6 * sizeof(double), cudaMemcpyDeviceToHost, st));
      LQG_CUDA_TRY(cudaStreamSynchronize(st));
      return LQG_OK;
    };
    // one Newton step of the reduced KKT system for the complementarity target chosen by `corrector`
    auto newton = [&](int corrector, double sigma_mu) -> int {
      LQG_CUDA_TRY(launch_qp_prepare(n, mc, corrector, sigma_mu, s, lam, rp, rd, ds, dl, rcv, t, rhs, st));
      LQG_CUDA_TRY(launch_gemm(true, false, n, 1, mc, 1.0, Ad, mc, t, mc, 1.0, rhs, n, 0, st));    // rhs += A' t
      LQG_CUDA_TRY(launch_chol_solve(n, Lc, n, rhs, n, 1, st));                        // dx
      LQG_CUDA_TRY(launch_gemm(false, false, mc, 1, n, 1.0, Ad, mc, rhs, n, 1.0, ds, mc, 0, st));   // ds = A dx + rp
      LQG_CUDA_TRY(launch_qp_step(mc, s, lam, ds, rcv, dl, scal, st));
      return read_scal();
    };
    LQG_CUDA_TRY(launch_gemm(false, false, mc, 1, n, 1.0, Ad, mc, x, n, 0.0, Ax, mc, 0, st));
    LQG_CUDA_TRY(launch_qp_init(mc, Ax, bd, s, lam, st));
    for (; it < max_iter; ++it) {
      // rd = D x - dvec - A' lam;  rp = A x - s - b;  mu = s.lam / mc
      LQG_CUDA_TRY(launch_qp_neg_copy(n, i_dv.dev, rd, st));
      LQG_CUDA_TRY(launch_gemm(false, false, n, 1, n, 1.0, i_D.dev, n, x, n, 1.0, rd, n, 0, st));
      LQG_CUDA_TRY(launch_gemm(true, false, n, 1, mc, -1.0, Ad, mc, lam, mc, 1.0, rd, n, 0, st));
      LQG_CUDA_TRY(launch_gemm(false, false, mc, 1, n, 1.0, Ad, mc, x, n, 0.0, Ax, mc, 0, st));
      LQG_CUDA_TRY(launch_qp_resid(n, mc, rd, Ax, s, bd, lam, rp, scal, st));
      if ((rc = read_scal())) return rc;
      const double mu = h[2] / mc;
      if (!(std::isfinite(h[0]) && std::isfinite(h[1]) && std::isfinite(mu))) {
        set_error("lqg_solve_qp: the iteration diverged (non-finite residuals at iteration %d)", it);
        fold_launches();
        return LQG_ERR_INVALID;
      }
      // merit phi: converged below tol; the best iterate is kept because close to the solution the
      // attainable dual residual grows like eps |A|^2 lam^2 / mu and phi can bottom out just above tol
      const double phi = std::max(std::max(h[0] / (1e3 * scale_d), h[1] / scale_p), mu / scale_p);
      if (phi < tol) { converged = true; break; }
      if (phi < best_phi) {
        best_phi = phi; stall = 0;
        LQG_CUDA_TRY(cudaMemcpyAsync(b_best.p, x, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
        LQG_CUDA_TRY(cudaMemcpyAsync(b_best.as<double>() + n, s, 2 * (size_t)mc * sizeof(double), cudaMemcpyDeviceToDevice, st));
      } else if (++stall >= 5) {
        break;
      }
      // H = D + A' diag(lam/s) A = Lc Lc'
      LQG_CUDA_TRY(launch_qp_scale_rows(mc, n, Ad, lam, s, AW, st));
      LQG_CUDA_TRY(cudaMemcpyAsync(H, i_D.dev, nn * sizeof(double), cudaMemcpyDeviceToDevice, st));
      LQG_CUDA_TRY(launch_gemm(true, false, n, n, mc, 1.0, Ad, mc, AW, mc, 1.0, H, n, 0, st));
      if ((rc = chol_device(n, H, 0, Lc, &bad, st))) return rc;
      if (bad) {                                                     // same rescue as the host restatement
        std::vector<double> diag(n);
        LQG_CUDA_TRY(cudaMemcpy2DAsync(diag.data(), sizeof(double), H, (size_t)(n + 1) * sizeof(double), sizeof(double), n,
                                       cudaMemcpyDeviceToHost, st));
        LQG_CUDA_TRY(cudaStreamSynchronize(st));
        double tr = 0.0;
        for (double v : diag) tr += v;
        LQG_CUDA_TRY(launch_add_diag(n, H, n, 1e-12 * tr / n, st));
        if ((rc = chol_device(n, H, 0, Lc, &bad, st))) return rc;
        if (bad) { set_error("lqg_solve_qp: the KKT matrix is not positive definite at iteration %d", it); fold_launches(); return LQG_ERR_NOT_PD; }
      }
      auto amax = [&]() { return std::min(1.0, std::min(h[3], h[4])); };
      if ((rc = newton(0, 0.0))) return rc;                          // predictor (affine scaling)
      const double a_aff = amax();
      LQG_CUDA_TRY(launch_qp_muaff(mc, a_aff, s, lam, ds, dl, scal, st));
      if ((rc = read_scal())) return rc;
      const double mu_aff = h[5] / mc;
      const double sigma = (mu_aff / mu) * (mu_aff / mu) * (mu_aff / mu);
      if ((rc = newton(1, sigma * mu))) return rc;                   // corrector
      const double a = 0.995 * amax();
      LQG_CUDA_TRY(launch_qp_update(n, mc, a, x, rhs, s, ds, lam, dl, st));
    }
    const double* s_fin = s; const double* lam_fin = lam;
    if (!converged && best_phi < 1e3 * tol) {                        // rounding floor reached: best iterate
      converged = true;
      x_result = b_best.as<double>();
      s_fin = b_best.as<double>() + n; lam_fin = s_fin + mc;
    }
    // ---- active-set polish: quadprog returns the exact minimiser, an interior-point iterate is only
    //      O(sqrt(mu)) close on weakly active rows.  Guess the active set (lam_i > s_i), solve the
    //      equality-constrained problem through the Schur complement (A_a D^-1 A_a') lam_a = b_a - A_a x0,
    //      x = x0 + D^-1 A_a' lam_a, check the KKT conditions of the full problem, adjust the set a few
    //      times.  Without a consistent set the interior-point iterate stands.
    if (converged && polish) {
      std::vector<double> h_s(mc), h_lam(mc), h_r(mc), h_la;
      LQG_CUDA_TRY(cudaMemcpyAsync(h_s.data(), s_fin, mc * sizeof(double), cudaMemcpyDeviceToHost, st));
      LQG_CUDA_TRY(cudaMemcpyAsync(h_lam.data(), lam_fin, mc * sizeof(double), cudaMemcpyDeviceToHost, st));
      LQG_CUDA_TRY(cudaStreamSynchronize(st));
      std::vector<char> act(mc);
      for (int i = 0; i < mc; ++i) act[i] = h_lam[i] > h_s[i];
      if ((rc = chol_device(n, i_D.dev, 0, Lc, &bad, st))) return rc;          // Lc held the last KKT factor
      DevBuf b_idx, b_AaT, b_Y, b_S, b_Ls, b_la, b_xp, b_rf;
      LQG_CUDA_TRY(b_xp.alloc(n * sizeof(double)));
      LQG_CUDA_TRY(b_rf.alloc(mc * sizeof(double)));
      const double* x0 = b_x0.as<double>();
      bool done = false;
      for (int round = 0; round < 8 && !done && !bad; ++round) {
        std::vector<int> idx;
        for (int i = 0; i < mc; ++i) if (act[i]) idx.push_back(i);
        const int k = (int)idx.size();
        double* xp = b_xp.as<double>();
        LQG_CUDA_TRY(cudaMemcpyAsync(xp, x0, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
        h_la.assign(k, 0.0);
        if (k > n) break;                                                      // more active rows than unknowns: dependent
        if (k > 0) {
          LQG_CUDA_TRY(b_idx.alloc(k * sizeof(int)));
          LQG_CUDA_TRY(b_AaT.alloc((size_t)n * k * sizeof(double)));
          LQG_CUDA_TRY(b_Y.alloc((size_t)n * k * sizeof(double)));
          LQG_CUDA_TRY(b_S.alloc((size_t)k * k * sizeof(double)));
          LQG_CUDA_TRY(b_Ls.alloc((size_t)k * k * sizeof(double)));
          LQG_CUDA_TRY(b_la.alloc(k * sizeof(double)));
          LQG_CUDA_TRY(cudaMemcpyAsync(b_idx.p, idx.data(), k * sizeof(int), cudaMemcpyHostToDevice, st));
          double* AaT = b_AaT.as<double>(); double* Y = b_Y.as<double>(); double* la = b_la.as<double>();
          LQG_CUDA_TRY(launch_qp_gather_rows(n, k, mc, Ad, b_idx.as<int>(), AaT, st));
          LQG_CUDA_TRY(cudaMemcpyAsync(Y, AaT, (size_t)n * k * sizeof(double), cudaMemcpyDeviceToDevice, st));
          LQG_CUDA_TRY(launch_chol_solve(n, Lc, n, Y, n, k, st));                                      // Y = D^-1 A_a'
          LQG_CUDA_TRY(launch_gemm(true, false, k, k, n, 1.0, AaT, n, Y, n, 0.0, b_S.as<double>(), k, 0, st));
          int bad_s = 0;
          if ((rc = chol_device(k, b_S.as<double>(), 0, b_Ls.as<double>(), &bad_s, st))) return rc;
          if (bad_s) break;                                                    // dependent active rows
          // la = b_a - A_a x0, then solve
          LQG_CUDA_TRY(launch_gemm(true, false, k, 1, n, 1.0, AaT, n, x0, n, 0.0, la, k, 0, st));
          LQG_CUDA_TRY(cudaMemcpyAsync(h_la.data(), la, k * sizeof(double), cudaMemcpyDeviceToHost, st));
          LQG_CUDA_TRY(cudaStreamSynchronize(st));
          for (int q = 0; q < k; ++q) h_la[q] = h_b[idx[q]] - h_la[q];
          LQG_CUDA_TRY(cudaMemcpyAsync(la, h_la.data(), k * sizeof(double), cudaMemcpyHostToDevice, st));
          LQG_CUDA_TRY(launch_chol_solve(k, b_Ls.as<double>(), k, la, k, 1, st));
          LQG_CUDA_TRY(launch_gemm(false, false, n, 1, k, 1.0, Y, n, la, k, 1.0, xp, n, 0, st));       // xp = x0 + Y la
          LQG_CUDA_TRY(cudaMemcpyAsync(h_la.data(), la, k * sizeof(double), cudaMemcpyDeviceToHost, st));
        }
        LQG_CUDA_TRY(launch_gemm(false, false, mc, 1, n, 1.0, Ad, mc, xp, n, 0.0, b_rf.as<double>(), mc, 0, st));
        LQG_CUDA_TRY(cudaMemcpyAsync(h_r.data(), b_rf.p, mc * sizeof(double), cudaMemcpyDeviceToHost, st));
        LQG_CUDA_TRY(cudaStreamSynchronize(st));
        double la_max = 1.0;
        for (double v : h_la) la_max = std::max(la_max, std::fabs(v));
        bool changed = false, ok = true;
        for (int q = 0; q < k; ++q)
          if (h_la[q] < -1e-10 * la_max) { act[idx[q]] = 0; changed = true; ok = false; }
        for (int i = 0; i < mc; ++i)
          if (h_r[i] - h_b[i] < -1e-12 * scale_p) { ok = false; if (!act[i]) { act[i] = 1; changed = true; } }
        if (ok) {                                                              // b_xp dies with this block: keep the result in b_best
          LQG_CUDA_TRY(cudaMemcpyAsync(b_best.p, xp, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
          x_result = b_best.as<double>();
          done = true;
        }
        else if (!changed) break;
      }
    }
  }
  if (iters) *iters = it;
  if ((rc = deliver(x_out, x_result, n, mem, st))) return rc;
  LQG_CUDA_TRY(cudaStreamSynchronize(st));
  fold_launches();
  if (!converged) {
    set_error("lqg_solve_qp: no convergence after %d iterations", it);
    return LQG_ERR_BUDGET;
  }
  return LQG_OK;
  LQG_API_END("lqg_solve_qp")
}

// mean and exact type-7 quantiles of a device-resident slab (ysim n_t x N, ld): bracketed one-pass
// path, exact transpose + radix-select path for the test points it flags or the shapes it does not cover
static int bands_on_device(int n_t, int64_t N, const double* y_dev, int64_t ld, const std::vector<double>& h_probs,
                           int nprobs, double* d_mean, double* d_q, DevBuf& b_scr, double* ms_out, cudaStream_t st) {
  static const char* env_exact = getenv("LQG_BANDS_EXACT");              // tuning / test aid
  const size_t fast_bytes = (env_exact && env_exact[0] == '1') ? 0 : bands_bracket_scratch_bytes(n_t, N, nprobs);
  const size_t need = fast_bytes ? fast_bytes : (size_t)n_t * N * sizeof(double);
  if (!b_scr.p || b_scr.bytes < need) LQG_CUDA_TRY(b_scr.alloc(need));
  Timer tm(st);
  tm.start();
  if (fast_bytes) {
    int* fail_dev = nullptr;
    LQG_CUDA_TRY(launch_bands_bracketed(n_t, N, y_dev, ld, h_probs.data(), nprobs, d_mean, d_q, b_scr.p, &fail_dev, st));
    std::vector<int> h_fail(n_t);
    LQG_CUDA_TRY(cudaMemcpyAsync(h_fail.data(), fail_dev, n_t * sizeof(int), cudaMemcpyDeviceToHost, st));
    LQG_CUDA_TRY(cudaStreamSynchronize(st));
    std::vector<int> rows;
    for (int r = 0; r < n_t; ++r) if (h_fail[r]) rows.push_back(r);
    if (getenv("LQG_TRACE")) fprintf(stderr, "[lqg bands] bracketed path: %zu of %d test points go to the exact path (first %d)\n",
                                     rows.size(), n_t, rows.empty() ? -1 : rows[0]);
    if (rows.size() > 64) {
      LQG_CUDA_TRY(b_scr.alloc((size_t)n_t * N * sizeof(double)));
      LQG_CUDA_TRY(launch_bands_exact(n_t, N, y_dev, ld, h_probs.data(), nprobs, d_mean, d_q, b_scr.as<double>(), st));
    } else if (!rows.empty()) {
      if (b_scr.bytes < (size_t)N * sizeof(double)) LQG_CUDA_TRY(b_scr.alloc((size_t)N * sizeof(double)));
      for (int r : rows)
        LQG_CUDA_TRY(launch_bands_exact(1, N, y_dev + r, ld, h_probs.data(), nprobs, d_mean + r,
                                        d_q + (size_t)r * nprobs, b_scr.as<double>(), st));
    }
  } else {
    LQG_CUDA_TRY(launch_bands_exact(n_t, N, y_dev, ld, h_probs.data(), nprobs, d_mean, d_q, b_scr.as<double>(), st));
  }
  tm.stop();
  if (ms_out) *ms_out = tm.ms();
  return LQG_OK;
}

// ===========================================================================================
static bool sparse_paths_enabled() {
  static const char* env = getenv("LQG_SPARSE");                     // A/B aid: 0 = always the dense DMMA products
  return !(env && env[0] == '0');
}
// Y (nr x N) = Phi_slab (nr x m) X (m x N).  A slab whose rows hold at most kEllMaxNnz non-zeros (1-D hat basis: 2)
// is multiplied in ELL form — bound by writing Y instead of an nr x m x N GEMM; the first dense slab ends the probing.
struct SlabProduct {
  DevBuf pidx, pval, pinfo;
  bool probe = sparse_paths_enabled();
  bool ell = false;                                                // of the slab classified last
  int ell_k = 0;
  int classify(int nr, int m, int rows_cap, const double* phi, int64_t ld, cudaStream_t st) {
    ell = false;
    ell_k = 0;
    if (!probe) return LQG_OK;
    if (!pidx.p) {
      LQG_CUDA_TRY(pidx.alloc((size_t)kEllMaxNnz * rows_cap * sizeof(int)));
      LQG_CUDA_TRY(pval.alloc((size_t)kEllMaxNnz * rows_cap * sizeof(double)));
      LQG_CUDA_TRY(pinfo.alloc(sizeof(int)));
    }
    LQG_CUDA_TRY(cudaMemsetAsync(pinfo.p, 0, sizeof(int), st));
    LQG_CUDA_TRY(launch_ell_from_dense(nr, m, kEllMaxNnz, phi, ld, pidx.as<int>(), pval.as<double>(), pinfo.as<int>(), st));
    LQG_CUDA_TRY(cudaMemcpyAsync(&ell_k, pinfo.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    LQG_CUDA_TRY(cudaStreamSynchronize(st));
    ell = ell_k <= kEllMaxNnz;
    if (!ell) probe = false;
    return LQG_OK;
  }
  int product(int nr, int m, const double* phi, int64_t ld, const double* X, int64_t ldx, int64_t N, double* Y, int64_t ldy,
              cudaStream_t st, double* ms) {
    Timer tg(st);
    tg.start();
    if (ell) LQG_CUDA_TRY(launch_ell_spmm(nr, m, std::max(1, ell_k), pidx.as<int>(), pval.as<double>(), X, ldx, N, Y, ldy, st));
    else LQG_CUDA_TRY(launch_dgemm(nr, (int)N, m, phi, ld, X, ldx, Y, ldy, nullptr, 0, st));
    tg.stop();
    if (ms) *ms += tg.ms();
    return LQG_OK;
  }
};

// Fusing pays when the band pass over a stored slab costs more than what the fused form adds to the product.  Per
// 4 224 x 125 208 slab on one B200: the band pass 1.9 ms (Gaussian values; 3.9 ms in the cfg5 pipeline), the fused
// epilogue + bracket sort 1.25 ms — both per value of the product, whatever m — and the pilot product 2048 / N of
// the product, which itself is 29.6 ms x m / 1000.  Hence: fuse iff (1.9 - 1.25) ms > (2048 / N) x 29.6 ms x m / 1000,
// i.e. 22 / m > 2048 / N (m = 1000: N > 93 000; measured break-even at N = 125 208: m between 1000 and 2000).
static bool bands_fused_enabled(int m, int64_t N) {
  static const char* env = getenv("LQG_BANDS_FUSED");                // A/B aid: 0 = always store the slab, 1 = fuse whenever possible
  if (env && env[0] == '0') return false;
  if (env && env[0] == '1') return true;
  return 22.0 / (double)m > 2048.0 / (double)N;
}

// One row slab of the path stage: the bands of Phi_slab X.  When the caller does not want the paths themselves and
// the slab is dense, the product is summarised in its own epilogue (launch_paths_bands_fused) and never stored;
// otherwise it goes into y (allocated on first use: rows_cap x N) and lqg_bands' kernels read it from there.
// *stored tells whether y holds the slab afterwards.
static int slab_paths_bands(SlabProduct& sp, int nr, int m, int rows_cap, const double* phi, int64_t ld, const double* X, int64_t ldx,
                            int64_t N, const std::vector<double>& h_probs, int nprobs, double* d_mean, double* d_q, DevBuf& b_y,
                            bool want_y, DevBuf& b_scr, DevBuf& b_row, double* t_gemm, double* t_bands, bool* stored, cudaStream_t st) {
  int rc;
  if ((rc = sp.classify(nr, m, rows_cap, phi, ld, st))) return rc;
  auto need_y = [&]() -> cudaError_t { return b_y.p ? cudaSuccess : b_y.alloc((size_t)rows_cap * N * sizeof(double)); };
  *stored = false;
  const size_t fused_bytes = (want_y || sp.ell || !bands_fused_enabled(m, N)) ? 0 : paths_bands_fused_scratch_bytes(nr, m, N, nprobs, phi, ld, X, ldx);
  if (!fused_bytes) {
    LQG_CUDA_TRY(need_y());
    if ((rc = sp.product(nr, m, phi, ld, X, ldx, N, b_y.as<double>(), nr, st, t_gemm))) return rc;
    *stored = true;
    double ms = 0.0;
    if ((rc = bands_on_device(nr, N, b_y.as<double>(), nr, h_probs, nprobs, d_mean, d_q, b_scr, &ms, st))) return rc;
    *t_bands += ms;
    return LQG_OK;
  }
  if (!b_scr.p || b_scr.bytes < fused_bytes) LQG_CUDA_TRY(b_scr.alloc(fused_bytes));
  Timer tg(st), tb(st);
  tg.start();
  int* fail_dev = nullptr;
  LQG_CUDA_TRY(launch_paths_bands_fused(nr, m, N, phi, ld, X, ldx, h_probs.data(), nprobs, d_mean, d_q, b_scr.p, &fail_dev, st, tg.b));
  std::vector<int> h_fail(nr);
  LQG_CUDA_TRY(cudaMemcpyAsync(h_fail.data(), fail_dev, nr * sizeof(int), cudaMemcpyDeviceToHost, st));
  LQG_CUDA_TRY(cudaStreamSynchronize(st));
  std::vector<int> rows;
  for (int r = 0; r < nr; ++r) if (h_fail[r]) rows.push_back(r);
  if (getenv("LQG_TRACE")) fprintf(stderr, "[lqg bands] fused product: %zu of %d test points go to the exact path (first %d)\n",
                                   rows.size(), nr, rows.empty() ? -1 : rows[0]);
  if (rows.size() > 64) {
    // the whole slab after all: same kernel arithmetic, stored this time
    LQG_CUDA_TRY(need_y());
    LQG_CUDA_TRY(launch_dgemm(nr, (int)N, m, phi, ld, X, ldx, b_y.as<double>(), nr, nullptr, 0, st));
    *stored = true;
    if (b_scr.bytes < (size_t)nr * N * sizeof(double)) LQG_CUDA_TRY(b_scr.alloc((size_t)nr * N * sizeof(double)));
    LQG_CUDA_TRY(launch_bands_exact(nr, N, b_y.as<double>(), nr, h_probs.data(), nprobs, d_mean, d_q, b_scr.as<double>(), st));
  } else if (!rows.empty()) {
    // single test points: their row of the product again (from an even row, so that the operand stays 16-byte
    // aligned and the same tensor-map kernel — the same rounding — produces it), then the exact selection
    if (!b_row.p || b_row.bytes < (size_t)3 * N * sizeof(double)) LQG_CUDA_TRY(b_row.alloc((size_t)3 * N * sizeof(double)));
    for (int r : rows) {
      const int rb = r & ~1, nrow = std::min(2, nr - rb);
      LQG_CUDA_TRY(launch_dgemm(nrow, (int)N, m, phi + rb, ld, X, ldx, b_row.as<double>(), nrow, nullptr, 0, st));
      LQG_CUDA_TRY(launch_bands_exact(1, N, b_row.as<double>() + (r - rb), nrow, h_probs.data(), nprobs, d_mean + r,
                                      d_q + (size_t)r * nprobs, b_row.as<double>() + (size_t)2 * N, st));
    }
  }
  tb.stop();
  float f = 0.f;
  *t_gemm += tg.ms();
  LQG_CUDA_TRY(cudaEventSynchronize(tb.b));
  if (cudaEventElapsedTime(&f, tg.b, tb.b) == cudaSuccess) *t_bands += f;
  return LQG_OK;
}

extern "C" int lqg_paths(int32_t n_t, int32_t m, int64_t N, const double* Phi, int64_t ldphi,
                         const double* xi, int64_t ldxi, const double* probs, int32_t nprobs,
                         double* ysim, int64_t ldy, double* mean, double* qtls, int32_t slab_rows,
                         const lqg_opts* opts_, double* gemm_ms, double* bands_ms) {
  LQG_API_BEGIN
  if (!have_device()) return LQG_ERR_CUDA;
  const lqg_opts o = opts_ ? *opts_ : default_opts();
  if (n_t <= 0 || m <= 0 || N <= 0 || N > 0x7fffffff || !Phi || !xi || ldphi < n_t || ldxi < m || nprobs < 0 || nprobs > 8 ||
      (nprobs > 0 && (!probs || !qtls)) || !mean || (ysim && ldy < n_t)) {
    set_error("lqg_paths: invalid argument (n_t=%d m=%d N=%lld nprobs=%d)", n_t, m, (long long)N, nprobs);
    return LQG_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)o.stream;
  DevBuf::cur_stream = st;
  const int mem = o.mem;
  int rc;
  std::vector<double> h_probs;
  if (nprobs > 0 && (rc = fetch_host(h_probs, probs, nprobs, mem, st))) return rc;
  for (int k = 0; k < nprobs; ++k)
    if (!(h_probs[k] >= 0.0 && h_probs[k] <= 1.0)) { set_error("'probs' outside [0,1]"); return LQG_ERR_INVALID; }
  Input i_phi, i_xi;
  if ((rc = i_phi.stage(Phi, (size_t)ldphi * (m - 1) + n_t, mem, st)) || (rc = i_xi.stage(xi, (size_t)ldxi * (N - 1) + m, mem, st)))
    return rc;
  int rows = slab_rows > 0 ? slab_rows : (int)std::max<int64_t>(1, std::min<int64_t>(n_t, ((int64_t)4 << 30) / (8 * N)));
  if (rows >= 128 && rows < n_t) rows -= rows % 128;              // whole 128-row GEMM tiles
  if (rows > n_t) rows = n_t;
  DevBuf b_y, b_scr, b_row, b_mean, b_q;
  double* d_mean = mean; double* d_q = qtls;
  if (mem == LQG_MEM_HOST) {
    LQG_CUDA_TRY(b_mean.alloc(n_t * sizeof(double))); d_mean = b_mean.as<double>();
    LQG_CUDA_TRY(b_q.alloc((size_t)std::max(1, nprobs) * n_t * sizeof(double))); d_q = b_q.as<double>();
  }
  double t_gemm = 0.0, t_bands = 0.0;
  SlabProduct slab;
  for (int r0 = 0; r0 < n_t; r0 += rows) {
    const int nr = std::min(rows, n_t - r0);
    bool stored = false;
    if ((rc = slab_paths_bands(slab, nr, m, rows, i_phi.dev + r0, ldphi, i_xi.dev, ldxi, N, h_probs, nprobs, d_mean + r0,
                               d_q + (size_t)r0 * nprobs, b_y, ysim != nullptr, b_scr, b_row, &t_gemm, &t_bands, &stored, st)))
      return rc;
    if (ysim) {
      if (mem == LQG_MEM_HOST) LQG_CUDA_TRY(d2h_result(ysim + r0, ldy * sizeof(double), b_y.p, (size_t)nr * sizeof(double), (size_t)nr * sizeof(double), N, st));
      else LQG_CUDA_TRY(cudaMemcpy2DAsync(ysim + r0, ldy * sizeof(double), b_y.p, (size_t)nr * sizeof(double), (size_t)nr * sizeof(double), N,
                                          cudaMemcpyDeviceToDevice, st));
      LQG_CUDA_TRY(cudaStreamSynchronize(st));                     // the slab buffer is reused by the next GEMM
    }
  }
  if (gemm_ms) *gemm_ms = t_gemm;
  if (bands_ms) *bands_ms = t_bands;
  if (mem == LQG_MEM_HOST) {
    LQG_CUDA_TRY(cudaMemcpyAsync(mean, d_mean, n_t * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (nprobs > 0) LQG_CUDA_TRY(cudaMemcpyAsync(qtls, d_q, (size_t)nprobs * n_t * sizeof(double), cudaMemcpyDeviceToHost, st));
  }
  LQG_CUDA_TRY(cudaStreamSynchronize(st));
  fold_launches();
  return LQG_OK;
  LQG_API_END("lqg_paths")
}

// ===========================================================================================
extern "C" int lqg_bands(int32_t n_t, int64_t N, const double* ysim, int64_t ld, const double* probs,
                         int32_t nprobs, double* mean, double* qtls, const lqg_opts* opts_,
                         double* kernel_ms) {
  LQG_API_BEGIN
  if (!have_device()) return LQG_ERR_CUDA;
  const lqg_opts o = opts_ ? *opts_ : default_opts();
  if (n_t <= 0 || N <= 0 || !ysim || ld < n_t || nprobs < 0 || nprobs > 8 || (nprobs > 0 && (!probs || !qtls)) || !mean) {
    set_error("lqg_bands: invalid argument (n_t=%d N=%lld nprobs=%d; at most 8 probabilities)", n_t, (long long)N, nprobs);
    return LQG_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)o.stream;
  DevBuf::cur_stream = st;
  const int mem = o.mem;
  int rc;
  std::vector<double> h_probs;
  if (nprobs > 0 && (rc = fetch_host(h_probs, probs, nprobs, mem, st))) return rc;
  for (int k = 0; k < nprobs; ++k)
    if (!(h_probs[k] >= 0.0 && h_probs[k] <= 1.0)) { set_error("'probs' outside [0,1]"); return LQG_ERR_INVALID; }
  Input i_y;
  if ((rc = i_y.stage(ysim, (size_t)ld * (N - 1) + n_t, mem, st))) return rc;
  DevBuf b_scr, b_mean, b_q;
  double* d_mean = mean; double* d_q = qtls;
  if (mem == LQG_MEM_HOST) {
    LQG_CUDA_TRY(b_mean.alloc(n_t * sizeof(double))); d_mean = b_mean.as<double>();
    LQG_CUDA_TRY(b_q.alloc((size_t)std::max(1, nprobs) * n_t * sizeof(double))); d_q = b_q.as<double>();
  }
  double ms = 0.0;
  if ((rc = bands_on_device(n_t, N, i_y.dev, ld, h_probs, nprobs, d_mean, d_q, b_scr, &ms, st))) return rc;
  if (kernel_ms) *kernel_ms = ms;
  if (mem == LQG_MEM_HOST) {
    LQG_CUDA_TRY(cudaMemcpyAsync(mean, d_mean, n_t * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (nprobs > 0) LQG_CUDA_TRY(cudaMemcpyAsync(qtls, d_q, (size_t)nprobs * n_t * sizeof(double), cudaMemcpyDeviceToHost, st));
  }
  LQG_CUDA_TRY(cudaStreamSynchronize(st));
  fold_launches();
  return LQG_OK;
  LQG_API_END("lqg_bands")
}

// ===========================================================================================
// lqg_simulate: sampler -> xi -> Phi xi -> bands in one call, device resident, optionally sharded
// over the GPUs of a communicator (include/lineqgpr_b200.h)
namespace {

// dst[:, g] = src[:, perm[g]]: puts a slab of sample paths, whose columns follow the gathered order of
// xi (round, rank, chain, sample), into job order (rank, chain, sample) = [xi_out(rank 0) | xi_out(rank 1) | ...]
__global__ void permute_cols_kernel(int rows, int64_t n_cols, const double* __restrict__ src, int64_t lds,
                                    const int* __restrict__ perm, double* __restrict__ dst, int64_t ldd) {
  const size_t total = (size_t)rows * n_cols;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t g = idx / rows;
    const int i = (int)(idx - g * rows);
    dst[g * ldd + i] = src[(size_t)perm[g] * lds + i];
  }
}

// Plan of xi = qr.solve(Lambda, eta) for a Lambda whose rows have their non-zeros in short windows (what
// lineqGPR builds): banded normal equations, see lqg_sparse.cu.  ok = false: use the dense Lambda^+ product.
struct LsqPlan {
  bool ok = false;
  LsqPlanDev dev{};
  DevBuf rstart, rval, cidx, cval, lband, fw, bw, info, gersh;
};
static int lsq_plan_build(int d, int m, const double* L_dev, LsqPlan& P, cudaStream_t st) {
  P.ok = false;
  if (!sparse_paths_enabled()) return LQG_OK;
  LQG_CUDA_TRY(P.rstart.alloc((size_t)d * sizeof(int)));
  LQG_CUDA_TRY(P.info.alloc(4 * sizeof(int)));
  LQG_CUDA_TRY(cudaMemsetAsync(P.info.p, 0, 4 * sizeof(int), st));
  LQG_CUDA_TRY(launch_lsq_profile(d, m, L_dev, P.rstart.as<int>(), P.info.as<int>(), st));
  int h[4] = {0, 0, 0, 0};
  LQG_CUDA_TRY(cudaMemcpyAsync(h, P.info.p, sizeof(h), cudaMemcpyDeviceToHost, st));
  LQG_CUDA_TRY(cudaStreamSynchronize(st));
  if (h[0] > kLsqMaxBand || h[1] > kLsqMaxColNnz || h[1] < 1) return LQG_OK;   // wide rows: dense path
  const int W = h[0] + 1, Wc = h[1], b = std::max(1, h[0]);
  LQG_CUDA_TRY(P.rval.alloc((size_t)W * d * sizeof(double)));
  LQG_CUDA_TRY(P.cidx.alloc((size_t)Wc * m * sizeof(int)));
  LQG_CUDA_TRY(P.cval.alloc((size_t)Wc * m * sizeof(double)));
  LQG_CUDA_TRY(P.lband.alloc((size_t)(b + 1) * m * sizeof(double)));
  LQG_CUDA_TRY(P.fw.alloc((size_t)(b + 1) * m * sizeof(double)));
  LQG_CUDA_TRY(P.bw.alloc((size_t)(b + 1) * m * sizeof(double)));
  LQG_CUDA_TRY(P.gersh.alloc(2 * sizeof(double)));
  LQG_CUDA_TRY(launch_lsq_build(d, m, W, Wc, b, L_dev, P.rstart.as<int>(), P.rval.as<double>(), P.cidx.as<int>(),
                                P.cval.as<double>(), P.lband.as<double>(), P.fw.as<double>(), P.bw.as<double>(),
                                P.info.as<int>(), P.gersh.as<double>(), st));
  double g[2] = {0.0, 0.0};
  LQG_CUDA_TRY(cudaMemcpyAsync(h, P.info.p, sizeof(h), cudaMemcpyDeviceToHost, st));
  LQG_CUDA_TRY(cudaMemcpyAsync(g, P.gersh.p, sizeof(g), cudaMemcpyDeviceToHost, st));
  LQG_CUDA_TRY(cudaStreamSynchronize(st));
  if (h[2] != 0) return LQG_OK;                                      // G not positive definite: let the dense path report it
  // normal equations lose cond(G) eps: with a proven cond(G) <= 1e3 that is below 1e-12 and the refinement
  // step (a second sparse product and solve) buys nothing
  const int refine = (g[0] > 0.0 && g[1] <= 1e3 * g[0]) ? 0 : 1;
  P.dev = LsqPlanDev{d, m, W, Wc, b, P.rstart.as<int>(), P.rval.as<double>(), P.cidx.as<int>(), P.cval.as<double>(),
                     P.fw.as<double>(), P.bw.as<double>(), refine};
  P.ok = true;
  return LQG_OK;
}

// 2-D device -> host copy of result columns on a copy stream: pinned destinations go through the
// copy engine directly, pageable ones (R / numpy matrices) through the staging ring
struct HostSink {
  cudaStream_t st = nullptr;
  std::unique_ptr<StagedCopier> staged;
  void open(cudaStream_t copy_st, const void* dst) {
    st = copy_st;
    static const char* env_staged = getenv("LQG_STAGED");
    if (!(env_staged && env_staged[0] == '0') && StagedCopier::is_pageable(dst)) staged.reset(new StagedCopier(copy_st));
  }
  cudaError_t copy2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t rows) {
    if (staged) return staged->copy2d(dst, dpitch, src, spitch, width, rows);
    return cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, cudaMemcpyDeviceToHost, st);
  }
  cudaError_t finish() { return staged ? staged->finish() : cudaSuccess; }
};

}  // namespace

extern "C" int lqg_simulate(const lqg_simulate_args* a_, const lqg_opts* opts_, lqg_comm* comm,
                            lqg_simulate_stats* stats) {
  LQG_API_BEGIN
  if (stats) memset(stats, 0, sizeof(*stats));
  if (!have_device()) return LQG_ERR_CUDA;
  if (!a_) { set_error("lqg_simulate: NULL arguments"); return LQG_ERR_INVALID; }
  const lqg_simulate_args a = *a_;
  const lqg_opts o = opts_ ? *opts_ : default_opts();
  const int d = a.d, m = a.m, n_t = a.n_t, nprobs = a.nprobs;
  const int n_chains = o.n_chains > 0 ? o.n_chains : 1;
  const int W = comm ? comm->world : 1, rank = comm ? comm->rank : 0;
  const bool have_pinv = a.Lambda_pinv != nullptr, have_lambda = a.Lambda != nullptr;
  if (d <= 0 || d > 8192 || m <= 0 || m > d || !a.mu || !a.Sigma || !a.lb || !a.ub || !a.init || a.n_per_chain <= 0 ||
      a.burn_in < 0 || (!have_pinv && !have_lambda && d != m) || n_t < 0 || nprobs < 0 || nprobs > 8 ||
      (nprobs > 0 && !a.probs) || (n_t > 0 && (!a.mean || (nprobs > 0 && !a.qtls))) ||
      (n_t > 0 && !a.Phi && (a.basis_D <= 0 || !a.basis_mk || !a.xtest || !a.knots || a.ldx < n_t)) ||
      (n_t > 0 && a.Phi && a.ldphi < n_t) || o.injected) {
    set_error("lqg_simulate: invalid argument (d=%d m=%d n_t=%d nprobs=%d n_per_chain=%d burn_in=%d%s)", d, m, n_t, nprobs,
              a.n_per_chain, a.burn_in, o.injected ? "; injected draws are a lqg_hmc_box feature" : "");
    return LQG_ERR_INVALID;
  }
  const int n_per = a.n_per_chain;
  const int64_t N_local = (int64_t)n_chains * n_per, N_total = N_local * W;
  if (N_total > 0x7fffffff) { set_error("lqg_simulate: %lld samples exceed the 2^31-1 columns of one job", (long long)N_total); return LQG_ERR_INVALID; }
  if (!a.Phi && n_t > 0) {
    long long mt = 0;
    for (int k = 0; k < a.basis_D; ++k) {
      if (a.basis_mk[k] < 2) { set_error("lqg_simulate: basis block %d has %d knots (needs >= 2)", k, a.basis_mk[k]); return LQG_ERR_INVALID; }
      mt += a.basis_mk[k];
    }
    if (mt != m) { set_error("lqg_simulate: the basis blocks have %lld knots in total, m = %d", mt, m); return LQG_ERR_INVALID; }
  }
  cudaStream_t st = (cudaStream_t)o.stream;
  DevBuf::cur_stream = st;
  const int mem = o.mem;
  const size_t dd = (size_t)d * d, dm = (size_t)d * m;
  int rc;
  Timer t_total(st);
  t_total.start();

  // ---- validation and staging of the box problem (as lqg_hmc_box) ----
  std::vector<double> h_lb, h_ub, h_init, h_probs;
  if ((rc = fetch_host(h_lb, a.lb, d, mem, st)) || (rc = fetch_host(h_ub, a.ub, d, mem, st)) ||
      (rc = fetch_host(h_init, a.init, d, mem, st)))
    return rc;
  if ((rc = validate_box(d, h_lb, h_ub, h_init, 0, 1))) return rc;
  if (nprobs > 0 && (rc = fetch_host(h_probs, a.probs, nprobs, mem, st))) return rc;
  for (int k = 0; k < nprobs; ++k)
    if (!(h_probs[k] >= 0.0 && h_probs[k] <= 1.0)) { set_error("'probs' outside [0,1]"); return LQG_ERR_INVALID; }
  Input i_mu, i_U, i_S, i_lb, i_ub, i_init, i_Lp;
  if ((rc = i_mu.stage(a.mu, d, mem, st)) || (a.U && (rc = i_U.stage(a.U, dd, mem, st))) ||
      (rc = i_S.stage(a.Sigma, dd, mem, st)) || (rc = i_lb.stage(a.lb, d, mem, st)) ||
      (rc = i_ub.stage(a.ub, d, mem, st)) || (rc = i_init.stage(a.init, d, mem, st)))
    return rc;
  int u_upper = 1;
  if (!a.U) {
    LQG_CUDA_TRY(i_U.buf.alloc(dd * sizeof(double)));
    if ((rc = ul_factor_device(d, i_S.dev, i_U.buf.as<double>(), st))) return rc;
    i_U.dev = i_U.buf.as<double>();
  }
  // ---- Lambda^+ (m x d) ----
  const bool identity = !have_pinv && !have_lambda;          // xi = eta
  // Lambda itself given: its rows are probed for short non-zero windows (identity / difference / block
  // diagonal constraint matrices); then xi comes from banded normal equations per sample (lqg_sparse.cu),
  // otherwise from the dense product with Lambda^+ (given, or computed here)
  LsqPlan lsq;
  Input i_L;
  if (have_lambda) {
    if ((rc = i_L.stage(a.Lambda, dm, mem, st))) return rc;
    if ((rc = lsq_plan_build(d, m, i_L.dev, lsq, st))) return rc;
  }
  if (!lsq.ok) {
    if (have_pinv) {
      if ((rc = i_Lp.stage(a.Lambda_pinv, dm, mem, st))) return rc;
    } else if (have_lambda) {
      LQG_CUDA_TRY(i_Lp.buf.alloc(dm * sizeof(double)));
      if ((rc = lambda_pinv_device(d, m, i_L.dev, i_Lp.buf.as<double>(), st))) return rc;   // synchronises st
      i_Lp.dev = i_Lp.buf.as<double>();
    }
  }

  // ---- device-resident eta and the gathered xi of the whole job ----
  DevBuf b_eta, b_xi;
  double* d_eta = (mem == LQG_MEM_DEVICE && a.eta_out) ? a.eta_out : nullptr;
  if (!d_eta) { LQG_CUDA_TRY(b_eta.alloc((size_t)d * N_local * sizeof(double))); d_eta = b_eta.as<double>(); }
  LQG_CUDA_TRY(b_xi.alloc((size_t)m * N_total * sizeof(double)));
  double* xi_all = b_xi.as<double>();

  StreamGuard side_guard, copy_guard;
  EventGuard ev_x_guard, ev_side_guard;
  {
    // the xi stage runs beside the sampling rounds, whose kernels fill every SM: with the highest priority its
    // (short) CTAs are placed as soon as anything retires instead of queueing behind a whole GEMM
    int prio_lo = 0, prio_hi = 0;
    LQG_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    LQG_CUDA_TRY(cudaStreamCreateWithPriority(&side_guard.s, cudaStreamNonBlocking, prio_hi));
  }
  LQG_CUDA_TRY(cudaEventCreateWithFlags(&ev_x_guard.e, cudaEventDisableTiming));
  LQG_CUDA_TRY(cudaEventCreateWithFlags(&ev_side_guard.e, cudaEventDisableTiming));
  cudaStream_t s_side = side_guard.s;
  HostSink sink_xi, sink_eta;
  const bool host_out = mem == LQG_MEM_HOST && (a.xi_out || a.eta_out);
  if (host_out) {
    LQG_CUDA_TRY(cudaStreamCreateWithFlags(&copy_guard.s, cudaStreamNonBlocking));
    if (a.xi_out) sink_xi.open(copy_guard.s, a.xi_out);
    if (a.eta_out) sink_eta.open(copy_guard.s, a.eta_out);
  }
  cudaStream_t s_copy = copy_guard.s;

  size_t round_off = 0;                   // doubles of xi_all filled by earlier rounds (all ranks)
  int rounds = 0;
  std::vector<std::pair<int, int>> round_span;                  // [from, upto) of every processed unit
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> xi_timers;   // per-round device time of the xi product
  struct EvList { std::vector<std::pair<cudaEvent_t, cudaEvent_t>>& v; ~EvList() { for (auto& p : v) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); } } } ev_list{xi_timers};

  BoxRun r;
  r.d = d; r.mu = i_mu.dev; r.U = i_U.dev; r.u_upper = u_upper; r.Sigma = i_S.dev; r.lb = i_lb.dev; r.ub = i_ub.dev;
  r.init = i_init.dev; r.init_ld = 0; r.n_chains = n_chains; r.n_per_chain = n_per; r.burn_in = a.burn_in;
  r.seed = o.seed; r.chain_offset = o.chain_offset + (int64_t)rank * n_chains; r.prepass = o.prepass;
  r.max_restarts = o.max_restarts; r.exact_search = o.hmc_search == 1; r.no_window = o.hmc_search == 2; r.fork = o.fork; r.fork_burn_in = o.fork_burn_in; r.out = d_eta; r.st = st;
  static const char* env_tail = getenv("LQG_SIM_SHORT_TAIL");          // A/B aid
  r.short_tail = env_tail ? env_tail[0] != '0' : (host_out || W > 1);  // the last round's copy / gather cannot hide behind sampling
  // The xi products, the all-gathers and the copies work on UNITS of `unit` samples per chain, not on
  // the sampling rounds themselves: how far a round gets depends on the chains, i.e. differs from
  // rank to rank, while the sequence of all-gather calls (and their sizes) must be the same on every
  // rank.  Unit u = samples [u*unit, (u+1)*unit) of every chain is processed as soon as the local
  // chains have all passed it; NCCL pairs unit u of one rank with unit u of the others whenever
  // each of them gets there.  The column order of the gathered xi is therefore fixed: unit, rank,
  // chain, sample.
  static const char* env_unit = getenv("LQG_GATHER_UNIT");            // tuning aid
  // one GPU: 16 — the part of the last unit's copy that no compute hides stays small (32: +2 % step time, 8:
  // launch-bound); several GPUs: 32 — the host's copy bandwidth is shared and fewer, larger copies and
  // all-gathers do better (8 GPUs: 8.9 vs 8.3 M samples/s end to end)
  const int unit = std::max(1, env_unit ? atoi(env_unit) : (W > 1 ? 32 : 16));
  const int n_units = (n_per + unit - 1) / unit;
  int next_unit = 0;
  DevBuf b_lsq_scr;                          // refinement right-hand sides of one unit (banded path)
  if (lsq.ok) LQG_CUDA_TRY(b_lsq_scr.alloc((size_t)m * n_chains * std::min(unit, n_per) * sizeof(double)));
  auto process_unit = [&](int from, int upto, cudaEvent_t ready) -> int {
    const int w = upto - from;
    const size_t cols = (size_t)n_chains * w;
    double* dst = xi_all + round_off + (size_t)rank * cols * m;
    LQG_CUDA_TRY(cudaStreamWaitEvent(s_side, ready, 0));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    LQG_CUDA_TRY(cudaEventCreate(&e0));
    LQG_CUDA_TRY(cudaEventCreate(&e1));
    xi_timers.emplace_back(e0, e1);
    LQG_CUDA_TRY(cudaEventRecord(e0, s_side));
    if (identity)      // xi = eta: gather the round's columns into the contiguous round block
      LQG_CUDA_TRY(cudaMemcpy2DAsync(dst, (size_t)w * m * sizeof(double), d_eta + (size_t)from * d, (size_t)n_per * d * sizeof(double),
                                     (size_t)w * m * sizeof(double), n_chains, cudaMemcpyDeviceToDevice, s_side));
    else if (lsq.ok)   // banded least squares per sample (HBM-bound, leaves the tensor pipe to the momentum GEMM)
      LQG_CUDA_TRY(launch_lsq_apply(lsq.dev, ColBlocks{d_eta + (size_t)from * d, d, w, (int64_t)n_per * d}, (int64_t)cols, dst,
                                    b_lsq_scr.as<double>(), s_side));
    else               // xi = Lambda^+ eta on the FP64 tensor cores; B = the round's column blocks of eta
      LQG_CUDA_TRY(launch_dgemm(m, (int)cols, d, i_Lp.dev, m, d_eta + (size_t)from * d, d, dst, m, nullptr, 0, s_side,
                                w, (int64_t)n_per * d));
    LQG_CUDA_TRY(cudaEventRecord(e1, s_side));
    LQG_CUDA_TRY(cudaEventRecord(ev_x_guard.e, s_side));
    int rcg;
    if (W > 1 && (rcg = comm_allgather_f64(comm, dst, xi_all + round_off, cols * m, s_side))) return rcg;
    const size_t spitch = (size_t)w * m * sizeof(double), dpitch = (size_t)n_per * m * sizeof(double);
    if (a.xi_out) {
      if (mem == LQG_MEM_HOST) {
        LQG_CUDA_TRY(cudaStreamWaitEvent(s_copy, ev_x_guard.e, 0));
        LQG_CUDA_TRY(sink_xi.copy2d(a.xi_out + (size_t)from * m, dpitch, dst, spitch, spitch, n_chains));
      } else {
        LQG_CUDA_TRY(cudaMemcpy2DAsync(a.xi_out + (size_t)from * m, dpitch, dst, spitch, spitch, n_chains, cudaMemcpyDeviceToDevice, s_side));
      }
    }
    if (a.eta_out && mem == LQG_MEM_HOST) {
      const size_t pe = (size_t)n_per * d * sizeof(double), we = (size_t)w * d * sizeof(double);
      LQG_CUDA_TRY(cudaStreamWaitEvent(s_copy, ready, 0));
      LQG_CUDA_TRY(sink_eta.copy2d(a.eta_out + (size_t)from * d, pe, d_eta + (size_t)from * d, pe, we, n_chains));
    }
    round_off += (size_t)W * cols * m;
    round_span.emplace_back(from, upto);
    ++rounds;
    return LQG_OK;
  };
  r.on_round = [&](int /*from*/, int upto, cudaEvent_t ready) -> int {
    while (next_unit < n_units) {
      const int u0 = next_unit * unit, u1 = std::min(n_per, u0 + unit);
      if (u1 > upto) break;
      const int rcu = process_unit(u0, u1, ready);
      if (rcu) return rcu;
      ++next_unit;
    }
    return LQG_OK;
  };
  BoxResult res;
  if ((rc = hmc_box_core(r, res))) return rc;
  // the path stage (caller's stream) starts when the last product / gather is done
  LQG_CUDA_TRY(cudaEventRecord(ev_side_guard.e, s_side));
  LQG_CUDA_TRY(cudaStreamWaitEvent(st, ev_side_guard.e, 0));
  int rc_chain = box_result_rc(res, "lqg_simulate");
  // a failed chain is an error on this rank, but the collectives below must still be entered by every
  // rank of the job: carry on and report at the end

  // ---- Phi xi slabs and their bands: rows [r0, r1) of Phi on this rank, all N_total samples ----
  double t_gemm = 0.0, t_bands = 0.0;
  const int rows_per = n_t > 0 ? (n_t + W - 1) / W : 0;
  const int r0 = std::min(n_t, rank * rows_per), r1 = std::min(n_t, r0 + rows_per);
  const int nr_local = r1 - r0;
  if (n_t > 0) {
    const int bw = 1 + nprobs;                                   // per test point: mean + nprobs quantiles
    DevBuf b_band, b_y, b_scr, b_row, b_phi, b_x, b_u;
    LQG_CUDA_TRY(b_band.alloc((size_t)W * rows_per * bw * sizeof(double)));
    LQG_CUDA_TRY(cudaMemsetAsync(b_band.p, 0, b_band.bytes, st));
    double* band_mine = b_band.as<double>() + (size_t)rank * rows_per * bw;   // [mean (rows_per) | qtls (nprobs x rows_per)]
    if (nr_local > 0) {
      int rows = a.slab_rows > 0 ? a.slab_rows : (int)std::max<int64_t>(1, std::min<int64_t>(nr_local, ((int64_t)4 << 30) / (8 * N_total)));
      if (rows >= 128 && rows < nr_local) rows -= rows % 128;    // whole 128-row GEMM tiles
      if (rows > nr_local) rows = nr_local;
      // this rank's rows of Phi: staged from the host (rows r0..r1 only), used in place on the device,
      // or built slab by slab from the test points
      const double* phi_dev = nullptr; int64_t phi_ld = 0;
      if (a.Phi) {
        if (mem == LQG_MEM_DEVICE) { phi_dev = a.Phi + r0; phi_ld = a.ldphi; }
        else {
          LQG_CUDA_TRY(b_phi.alloc((size_t)nr_local * m * sizeof(double)));
          LQG_CUDA_TRY(cudaMemcpy2DAsync(b_phi.p, (size_t)nr_local * sizeof(double), a.Phi + r0, (size_t)a.ldphi * sizeof(double),
                                         (size_t)nr_local * sizeof(double), m, cudaMemcpyHostToDevice, st));
          phi_dev = b_phi.as<double>(); phi_ld = nr_local;
        }
      } else {
        LQG_CUDA_TRY(b_phi.alloc((size_t)rows * m * sizeof(double)));
        if (mem == LQG_MEM_HOST) {
          LQG_CUDA_TRY(b_x.alloc((size_t)nr_local * a.basis_D * sizeof(double)));
          LQG_CUDA_TRY(cudaMemcpy2DAsync(b_x.p, (size_t)nr_local * sizeof(double), a.xtest + r0, (size_t)a.ldx * sizeof(double),
                                         (size_t)nr_local * sizeof(double), a.basis_D, cudaMemcpyHostToDevice, st));
          LQG_CUDA_TRY(b_u.alloc((size_t)m * sizeof(double)));
          LQG_CUDA_TRY(cudaMemcpyAsync(b_u.p, a.knots, (size_t)m * sizeof(double), cudaMemcpyHostToDevice, st));
        }
      }
      DevBuf b_flag, b_perm, b_yp;
      LQG_CUDA_TRY(b_flag.alloc(sizeof(int)));
      LQG_CUDA_TRY(cudaMemsetAsync(b_flag.p, 0, sizeof(int), st));
      if (a.ysim_out) {
        if (a.ldy < nr_local) { set_error("lqg_simulate: ldy=%lld < %d rows of this rank", (long long)a.ldy, nr_local); return LQG_ERR_INVALID; }
        // job-order column g = (rank, chain, sample)  ->  its column in the gathered order
        std::vector<int> perm((size_t)N_total);
        size_t col_off = 0;
        for (const auto& sp : round_span) {
          const int w = sp.second - sp.first;
          for (int rk = 0; rk < W; ++rk)
            for (int c = 0; c < n_chains; ++c)
              for (int s = sp.first; s < sp.second; ++s)
                perm[(size_t)rk * N_local + (size_t)c * n_per + s] = (int)(col_off + ((size_t)rk * n_chains + c) * w + (s - sp.first));
          col_off += (size_t)W * n_chains * w;
        }
        LQG_CUDA_TRY(b_perm.alloc(perm.size() * sizeof(int)));
        LQG_CUDA_TRY(cudaMemcpyAsync(b_perm.p, perm.data(), perm.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        LQG_CUDA_TRY(cudaStreamSynchronize(st));               // perm is a local
        if (mem == LQG_MEM_HOST) LQG_CUDA_TRY(b_yp.alloc((size_t)rows * N_total * sizeof(double)));
      }
      SlabProduct slab;
      for (int s0 = 0; s0 < nr_local; s0 += rows) {
        const int nr = std::min(rows, nr_local - s0);
        const double* phi_slab; int64_t ld_slab;
        if (a.Phi) { phi_slab = phi_dev + s0; ld_slab = phi_ld; }
        else {
          const double* x_dev = mem == LQG_MEM_HOST ? b_x.as<double>() + s0 : a.xtest + r0 + s0;
          const int64_t x_ld = mem == LQG_MEM_HOST ? nr_local : a.ldx;
          const double* u_dev = mem == LQG_MEM_HOST ? b_u.as<double>() : a.knots;
          size_t off = 0;
          for (int k = 0; k < a.basis_D; ++k) {
            LQG_CUDA_TRY(launch_basis_hat(nr, a.basis_mk[k], x_dev + (size_t)k * x_ld, u_dev + off, b_phi.as<double>() + off * nr, nr,
                                          b_flag.as<int>(), st));
            off += (size_t)a.basis_mk[k];
          }
          phi_slab = b_phi.as<double>(); ld_slab = nr;
        }
        double* q_slab = nprobs > 0 ? band_mine + rows_per + (size_t)s0 * nprobs : nullptr;
        // the slab's quantiles land in a (nprobs x rows_per) column-major block: column = test point
        bool stored = false;
        if ((rc = slab_paths_bands(slab, nr, m, rows, phi_slab, ld_slab, xi_all, m, N_total, h_probs, nprobs, band_mine + s0,
                                   nprobs > 0 ? q_slab : band_mine + rows_per, b_y, a.ysim_out != nullptr, b_scr, b_row, &t_gemm,
                                   &t_bands, &stored, st)))
          return rc;
        if (a.ysim_out) {
          double* yp = mem == LQG_MEM_HOST ? b_yp.as<double>() : a.ysim_out + s0;
          const int64_t ldp = mem == LQG_MEM_HOST ? nr : a.ldy;
          permute_cols_kernel<<<148 * 8, 256, 0, st>>>(nr, N_total, b_y.as<double>(), nr, b_perm.as<int>(), yp, ldp);
          ++launch_counter();
          LQG_CUDA_TRY(cudaGetLastError());
          if (mem == LQG_MEM_HOST)
            LQG_CUDA_TRY(d2h_result(a.ysim_out + s0, a.ldy * sizeof(double), yp, (size_t)nr * sizeof(double), (size_t)nr * sizeof(double), N_total, st));
          LQG_CUDA_TRY(cudaStreamSynchronize(st));                 // the slab buffers are reused by the next GEMM
        }
      }
      int flag = 0;
      LQG_CUDA_TRY(cudaMemcpyAsync(&flag, b_flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
      LQG_CUDA_TRY(cudaStreamSynchronize(st));
      if (flag && rc_chain == LQG_OK) { set_error("basisCompute.lineqGP: x outside the knot range [u_1, u_m]"); rc_chain = LQG_ERR_INVALID; }
    }
    // ---- all-gather of the band pieces, unpacked on the host ----
    if (W > 1 && (rc = comm_allgather_f64(comm, band_mine, b_band.as<double>(), (size_t)rows_per * bw, st))) return rc;
    std::vector<double> h_band((size_t)W * rows_per * bw);
    LQG_CUDA_TRY(cudaMemcpyAsync(h_band.data(), b_band.p, h_band.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    LQG_CUDA_TRY(cudaStreamSynchronize(st));
    std::vector<double> h_mean(n_t), h_q((size_t)std::max(1, nprobs) * n_t);
    for (int rk = 0; rk < W; ++rk) {
      const int b0 = std::min(n_t, rk * rows_per), b1 = std::min(n_t, b0 + rows_per);
      const double* blk = h_band.data() + (size_t)rk * rows_per * bw;
      for (int i = b0; i < b1; ++i) {
        h_mean[i] = blk[i - b0];
        for (int k = 0; k < nprobs; ++k) h_q[(size_t)i * nprobs + k] = blk[rows_per + (size_t)(i - b0) * nprobs + k];
      }
    }
    if (mem == LQG_MEM_HOST) {
      memcpy(a.mean, h_mean.data(), n_t * sizeof(double));
      if (nprobs > 0) memcpy(a.qtls, h_q.data(), (size_t)nprobs * n_t * sizeof(double));
    } else {
      LQG_CUDA_TRY(cudaMemcpyAsync(a.mean, h_mean.data(), n_t * sizeof(double), cudaMemcpyHostToDevice, st));
      if (nprobs > 0) LQG_CUDA_TRY(cudaMemcpyAsync(a.qtls, h_q.data(), (size_t)nprobs * n_t * sizeof(double), cudaMemcpyHostToDevice, st));
      LQG_CUDA_TRY(cudaStreamSynchronize(st));
    }
  }
  // ---- outputs still in flight, timings ----
  if (s_copy) {
    if (a.xi_out) LQG_CUDA_TRY(sink_xi.finish());
    if (a.eta_out) LQG_CUDA_TRY(sink_eta.finish());
    LQG_CUDA_TRY(cudaStreamSynchronize(s_copy));
  }
  LQG_CUDA_TRY(cudaStreamSynchronize(s_side));
  t_total.stop();
  double xi_ms = 0.0;
  for (auto& pr : xi_timers) { float f = 0; if (cudaEventElapsedTime(&f, pr.first, pr.second) == cudaSuccess) xi_ms += f; }
  if (stats) {
    fill_stats(&stats->hmc, res.stats, res.chain_ms, res.pre_ms, res.stats[10], res.stats[11]);
    fill_root_stats(&stats->hmc, res.root_stats, res.root_chain_ms, res.root_pre_ms);
    stats->xi_ms = xi_ms; stats->gemm_ms = t_gemm; stats->bands_ms = t_bands; stats->total_ms = t_total.ms();
    stats->n_local = N_local; stats->n_total = N_total; stats->row_begin = r0; stats->row_end = r1;
    stats->rank = rank; stats->world = W; stats->rounds = rounds;
  }
  LQG_CUDA_TRY(cudaStreamSynchronize(st));
  fold_launches();
  return rc_chain;
  LQG_API_END("lqg_simulate")
}
