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

// pinned ring.  One process-wide instance is allocated on first use and kept; a copier that finds
// it busy (another host thread is inside the library) works on a private ring of its own, so
// concurrent callers do not serialise behind each other's sampling loops.
struct Ring {
  std::mutex owner;                       // held by a StagedCopier only while it is moving bytes
  unsigned char* base = nullptr;
  cudaEvent_t ev[kSlots] = {};
  bool ready = false;                     // base AND every event exist
  // called with `owner` held.  A partial failure releases what was created, so that a later call
  // starts from scratch instead of running with null events.
  cudaError_t init() {
    if (ready) return cudaSuccess;
    cudaError_t e = cudaHostAlloc((void**)&base, kSlotBytes * kSlots, cudaHostAllocDefault);
    if (e != cudaSuccess) { base = nullptr; return e; }
    for (int i = 0; i < kSlots; ++i)
      if ((e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming | cudaEventBlockingSync)) != cudaSuccess) {
        for (int j = 0; j < i; ++j) { cudaEventDestroy(ev[j]); ev[j] = nullptr; }
        cudaFreeHost(base);
        base = nullptr;
        return e;
      }
    ready = true;
    return cudaSuccess;
  }
  void destroy() {
    if (!ready) return;
    for (int i = 0; i < kSlots; ++i) { cudaEventDestroy(ev[i]); ev[i] = nullptr; }
    cudaFreeHost(base);
    base = nullptr;
    ready = false;
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
  std::unique_lock<std::mutex> own;       // engaged when `ring` is the process-wide ring
  Ring* ring = nullptr;
  Ring* private_ring = nullptr;
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
      cudaError_t e = cudaEventSynchronize(ring->ev[t.slot]);
      if (e == cudaSuccess) {
        const unsigned char* src = ring->base + (size_t)t.slot * kSlotBytes;
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
  impl_->own = std::unique_lock<std::mutex>(g_ring.owner, std::try_to_lock);
  if (impl_->own.owns_lock()) {
    impl_->ring = &g_ring;
  } else {
    impl_->private_ring = new Ring;
    impl_->ring = impl_->private_ring;
  }
  impl_->err = impl_->ring->init();
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
  if (impl_->private_ring) { impl_->private_ring->destroy(); delete impl_->private_ring; }
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
    unsigned char* stage = impl_->ring->base + (size_t)slot * kSlotBytes;
    cudaError_t e = (r == 1) ? cudaMemcpyAsync(stage, s, w, cudaMemcpyDeviceToHost, impl_->st)
                             : cudaMemcpy2DAsync(stage, w, s, sp, w, r, cudaMemcpyDeviceToHost, impl_->st);
    if (e == cudaSuccess) e = cudaEventRecord(impl_->ring->ev[slot], impl_->st);
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
