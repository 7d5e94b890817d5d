// Multi-GPU exchange step of the simulate pipeline (option A): one process per GPU,
// chains sharded with no data-path collective, and exactly two exchanges at the end —
// an all-gather of the xi shards (so that every rank can summarise ITS rows of Phi over ALL samples)
// and an all-gather of the band pieces.  Both are NCCL all-gathers, device to device over NVLink.
//
// NCCL is resolved at run time (dlopen): the library has no link-time dependency on it, loads on a
// box without NCCL, and inside a torch process picks up the libnccl that torch already mapped
// instead of bringing a second copy.  The communicator is created from a 128-byte unique id that
// the host program distributes however it likes (torch.distributed, MPI, a file) — the C ABI takes
// bytes, not torch types.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "lqg_common.cuh"
#include "lqg_comm.h"

namespace lqg {

namespace {

// the handful of NCCL declarations this file needs (nccl.h: ncclUniqueId is 128 opaque bytes,
// ncclFloat64 = 8, ncclSuccess = 0)
struct NcclId { char internal[128]; };
typedef void* NcclComm;
constexpr int kNcclFloat64 = 8;

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
  char why[256] = "";
};
NcclApi g_api;
std::once_flag g_once;

void load_nccl() {
  const char* env = getenv("LQG_NCCL_LIB");
  const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
  // a copy that is already mapped into the process (torch's) wins over a fresh load
  for (const char* n : names) {
    if (!n || g_api.handle) continue;
    g_api.handle = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
  }
  for (const char* n : names) {
    if (!n || g_api.handle) continue;
    g_api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
  }
  if (!g_api.handle) {
    snprintf(g_api.why, sizeof(g_api.why), "libnccl.so.2 not found (%s); set LQG_NCCL_LIB", dlerror());
    return;
  }
  auto sym = [&](const char* s) { return dlsym(g_api.handle, s); };
  g_api.GetUniqueId = reinterpret_cast<int (*)(NcclId*)>(sym("ncclGetUniqueId"));
  g_api.CommInitRank = reinterpret_cast<int (*)(NcclComm*, int, NcclId, int)>(sym("ncclCommInitRank"));
  g_api.CommDestroy = reinterpret_cast<int (*)(NcclComm)>(sym("ncclCommDestroy"));
  g_api.AllGather = reinterpret_cast<int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t)>(sym("ncclAllGather"));
  g_api.GetErrorString = reinterpret_cast<const char* (*)(int)>(sym("ncclGetErrorString"));
  g_api.GetVersion = reinterpret_cast<int (*)(int*)>(sym("ncclGetVersion"));
  if (!g_api.GetUniqueId || !g_api.CommInitRank || !g_api.CommDestroy || !g_api.AllGather) {
    snprintf(g_api.why, sizeof(g_api.why), "the NCCL library lacks a required symbol");
    g_api.handle = nullptr;
  }
}

const NcclApi* api() {
  std::call_once(g_once, load_nccl);
  if (!g_api.handle) { set_error("NCCL unavailable: %s", g_api.why); return nullptr; }
  return &g_api;
}

int nccl_fail(const NcclApi* a, const char* what, int rc) {
  set_error("%s failed: %s", what, a->GetErrorString ? a->GetErrorString(rc) : "NCCL error");
  return LQG_ERR_CUDA;
}

}  // namespace

int comm_allgather_f64(const lqg_comm* c, const double* send, double* recv, size_t count, cudaStream_t st) {
  if (!c || c->world == 1) {
    if (send != recv)
      LQG_CUDA_TRY(cudaMemcpyAsync(recv, send, count * sizeof(double), cudaMemcpyDeviceToDevice, st));
    return LQG_OK;
  }
  const NcclApi* a = api();
  if (!a) return LQG_ERR_CUDA;
  const int rc = a->AllGather(send, recv, count, kNcclFloat64, (NcclComm)c->nccl, st);
  if (rc != 0) return nccl_fail(a, "ncclAllGather", rc);
  ++launch_counter();
  return LQG_OK;
}

}  // namespace lqg

using namespace lqg;

extern "C" int lqg_comm_unique_id(void* id128) {
  if (!id128) { set_error("lqg_comm_unique_id: NULL buffer"); return LQG_ERR_INVALID; }
  const NcclApi* a = api();
  if (!a) return LQG_ERR_CUDA;
  NcclId id;
  const int rc = a->GetUniqueId(&id);
  if (rc != 0) return nccl_fail(a, "ncclGetUniqueId", rc);
  memcpy(id128, id.internal, sizeof(id.internal));
  return LQG_OK;
}

extern "C" int lqg_comm_init(const void* id128, int32_t rank, int32_t world, lqg_comm** out) {
  if (!out) { set_error("lqg_comm_init: NULL result pointer"); return LQG_ERR_INVALID; }
  *out = nullptr;
  if (world <= 0 || rank < 0 || rank >= world || (world > 1 && !id128)) {
    set_error("lqg_comm_init: invalid argument (rank=%d, world=%d)", rank, world);
    return LQG_ERR_INVALID;
  }
  lqg_comm* c = new (std::nothrow) lqg_comm;
  if (!c) { set_error("lqg_comm_init: out of memory"); return LQG_ERR_INTERNAL; }
  c->nccl = nullptr; c->rank = rank; c->world = world;
  if (world > 1) {
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
      cudaGetLastError();
      delete c;
      set_error("lqg_comm_init: no CUDA device available; lineqgpr_b200 has no CPU fallback");
      return LQG_ERR_CUDA;
    }
    const NcclApi* a = api();
    if (!a) { delete c; return LQG_ERR_CUDA; }
    NcclId id;
    memcpy(id.internal, id128, sizeof(id.internal));
    NcclComm nc = nullptr;
    const int rc = a->CommInitRank(&nc, world, id, rank);      // collective: every rank calls it
    if (rc != 0) { delete c; return nccl_fail(a, "ncclCommInitRank", rc); }
    c->nccl = nc;
  }
  *out = c;
  return LQG_OK;
}

extern "C" int lqg_comm_rank(const lqg_comm* c) { return c ? c->rank : 0; }
extern "C" int lqg_comm_world(const lqg_comm* c) { return c ? c->world : 1; }

extern "C" int lqg_comm_destroy(lqg_comm* c) {
  if (!c) return LQG_OK;
  int rc = LQG_OK;
  if (c->nccl) {
    const NcclApi* a = api();
    if (a) { const int r = a->CommDestroy((NcclComm)c->nccl); if (r != 0) rc = nccl_fail(a, "ncclCommDestroy", r); }
  }
  delete c;
  return rc;
}

// all-gather of `count` doubles per rank (device pointers; recv holds world * count doubles and may
// alias send at recv + rank * count).  Exposed so that host programs and tests can move shards with
// the library's communicator; lqg_simulate uses the same routine internally.
extern "C" int lqg_comm_allgather(const lqg_comm* c, const double* send, double* recv, int64_t count, void* stream) {
  if (!send || !recv || count < 0) { set_error("lqg_comm_allgather: invalid argument"); return LQG_ERR_INVALID; }
  return comm_allgather_f64(c, send, recv, (size_t)count, (cudaStream_t)stream);
}
