// Internal view of the communicator handle of the C ABI (include/lineqgpr_b200.h: lqg_comm).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

struct lqg_comm {
  void* nccl;        // ncclComm_t, NULL when world == 1
  int rank, world;
};

namespace lqg {
// all-gather of `count` doubles per rank on `st`; recv may alias send at recv + rank * count.
// world == 1 (or c == NULL) degenerates to a device copy.
int comm_allgather_f64(const lqg_comm* c, const double* send, double* recv, size_t count, cudaStream_t st);
}  // namespace lqg
