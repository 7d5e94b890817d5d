// Device -> pageable host copies through a pinned staging ring and helper threads.
//
// tmvrnorm returns a d x nsim matrix to its caller (R/lineqGPsamplers.R:142); in R (and numpy)
// that matrix is ordinary pageable memory.  A cudaMemcpy into pageable memory is staged by the
// driver through one internal buffer and one thread (measured ~6 GB/s: 0.33 s for the 2 GB of
// 125 000 samples at d = 2000, five times the sampling itself).  Here the copy engine writes
// into a ring of pinned slots at PCIe speed while a few helper threads move finished slots into
// the caller's pages in parallel (page faults included), overlapping the sampling rounds.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace lqg {

class StagedCopier {
 public:
  // true when `p` is ordinary (unregistered) host memory, i.e. when staging pays off
  static bool is_pageable(const void* p);
  explicit StagedCopier(cudaStream_t copy_stream);
  ~StagedCopier();                       // waits for outstanding work
  StagedCopier(const StagedCopier&) = delete;
  StagedCopier& operator=(const StagedCopier&) = delete;
  // rows x width bytes from device (src, pitch spitch) to host (dst, pitch dpitch); ordered after
  // everything already enqueued on copy_stream (the caller makes copy_stream wait for the producer)
  cudaError_t copy2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t rows);
  cudaError_t finish();                  // blocks until every byte is in place
 private:
  struct Impl;
  Impl* impl_;
};

}  // namespace lqg
