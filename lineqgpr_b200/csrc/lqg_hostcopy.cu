#include "lqg_hostcopy.h"

#include <string.h>

#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace lqg {

namespace {

constexpr size_t kSlotBytes = (size_t)16 << 20;
constexpr int kSlots = 8;
constexpr int kWorkers = 4;

// process-wide pinned ring (allocated on first use, kept): one user at a time
struct Ring {
  std::mutex owner;                       // held by the StagedCopier that is using the ring
  unsigned char* base = nullptr;
  cudaEvent_t ev[kSlots] = {};
  cudaError_t init() {
    if (base) return cudaSuccess;
    cudaError_t e = cudaHostAlloc((void**)&base, kSlotBytes * kSlots, cudaHostAllocDefault);
    if (e != cudaSuccess) { base = nullptr; return e; }
    for (int i = 0; i < kSlots; ++i)
      if ((e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming | cudaEventBlockingSync)) != cudaSuccess) return e;
    return cudaSuccess;
  }
};
Ring g_ring;

struct Task {
  int slot;
  unsigned char* dst; size_t dpitch, width, rows;
};

}  // namespace

struct StagedCopier::Impl {
  cudaStream_t st;
  int device = 0;
  std::unique_lock<std::mutex> own;
  std::mutex m;
  std::condition_variable cv_task, cv_free;
  std::deque<Task> tasks;
  std::vector<int> free_slots;
  int in_flight = 0;
  bool stop = false;
  cudaError_t err = cudaSuccess;
  std::vector<std::thread> workers;

  void worker() {
    cudaSetDevice(device);
    for (;;) {
      Task t;
      {
        std::unique_lock<std::mutex> lk(m);
        cv_task.wait(lk, [&] { return stop || !tasks.empty(); });
        if (tasks.empty()) return;
        t = tasks.front();
        tasks.pop_front();
      }
      cudaError_t e = cudaEventSynchronize(g_ring.ev[t.slot]);
      if (e == cudaSuccess) {
        const unsigned char* src = g_ring.base + (size_t)t.slot * kSlotBytes;
        if (t.dpitch == t.width) memcpy(t.dst, src, t.width * t.rows);
        else for (size_t r = 0; r < t.rows; ++r) memcpy(t.dst + r * t.dpitch, src + r * t.width, t.width);
      }
      {
        std::lock_guard<std::mutex> lk(m);
        if (e != cudaSuccess && err == cudaSuccess) err = e;
        free_slots.push_back(t.slot);
        --in_flight;
      }
      cv_free.notify_all();
    }
  }
};

bool StagedCopier::is_pageable(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
  return a.type == cudaMemoryTypeUnregistered;
}

StagedCopier::StagedCopier(cudaStream_t copy_stream) : impl_(new Impl) {
  impl_->st = copy_stream;
  cudaGetDevice(&impl_->device);
  impl_->own = std::unique_lock<std::mutex>(g_ring.owner);
  impl_->err = g_ring.init();
  for (int i = 0; i < kSlots; ++i) impl_->free_slots.push_back(i);
  if (impl_->err == cudaSuccess)
    for (int i = 0; i < kWorkers; ++i) impl_->workers.emplace_back([this] { impl_->worker(); });
}

StagedCopier::~StagedCopier() {
  finish();
  {
    std::lock_guard<std::mutex> lk(impl_->m);
    impl_->stop = true;
  }
  impl_->cv_task.notify_all();
  for (auto& t : impl_->workers) t.join();
  delete impl_;
}

cudaError_t StagedCopier::copy2d(void* dst_, size_t dpitch, const void* src_, size_t spitch, size_t width, size_t rows) {
  if (impl_->err != cudaSuccess) return impl_->err;
  unsigned char* dst = static_cast<unsigned char*>(dst_);
  const unsigned char* src = static_cast<const unsigned char*>(src_);
  auto submit = [&](unsigned char* d, size_t dp, const unsigned char* s, size_t sp, size_t w, size_t r) -> cudaError_t {
    int slot;
    {
      std::unique_lock<std::mutex> lk(impl_->m);
      impl_->cv_free.wait(lk, [&] { return !impl_->free_slots.empty(); });
      slot = impl_->free_slots.back();
      impl_->free_slots.pop_back();
      ++impl_->in_flight;
    }
    unsigned char* stage = g_ring.base + (size_t)slot * kSlotBytes;
    cudaError_t e = (r == 1) ? cudaMemcpyAsync(stage, s, w, cudaMemcpyDeviceToHost, impl_->st)
                             : cudaMemcpy2DAsync(stage, w, s, sp, w, r, cudaMemcpyDeviceToHost, impl_->st);
    if (e == cudaSuccess) e = cudaEventRecord(g_ring.ev[slot], impl_->st);
    if (e != cudaSuccess) {
      std::lock_guard<std::mutex> lk(impl_->m);
      impl_->free_slots.push_back(slot);
      --impl_->in_flight;
      if (impl_->err == cudaSuccess) impl_->err = e;
      return e;
    }
    {
      std::lock_guard<std::mutex> lk(impl_->m);
      impl_->tasks.push_back(Task{slot, d, dp, w, r});
    }
    impl_->cv_task.notify_one();
    return cudaSuccess;
  };
  cudaError_t e;
  if (width <= kSlotBytes) {
    const size_t per = kSlotBytes / width;                    // rows per slot
    for (size_t r = 0; r < rows; r += per) {
      const size_t rc = rows - r < per ? rows - r : per;
      if ((e = submit(dst + r * dpitch, dpitch, src + r * spitch, spitch, width, rc)) != cudaSuccess) return e;
    }
  } else {
    for (size_t r = 0; r < rows; ++r)
      for (size_t off = 0; off < width; off += kSlotBytes) {
        const size_t w = width - off < kSlotBytes ? width - off : kSlotBytes;
        if ((e = submit(dst + r * dpitch + off, w, src + r * spitch + off, w, w, 1)) != cudaSuccess) return e;
      }
  }
  return cudaSuccess;
}

cudaError_t StagedCopier::finish() {
  std::unique_lock<std::mutex> lk(impl_->m);
  impl_->cv_free.wait(lk, [&] { return impl_->in_flight == 0; });
  return impl_->err;
}

}  // namespace lqg
