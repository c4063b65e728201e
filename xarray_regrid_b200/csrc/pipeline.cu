// C ABI of the persistent streaming pipeline + the threaded host copy used to stage pageable memory.
#include <cstring>
#include <new>
#include <thread>
#include <vector>

#include <pthread.h>

#include <atomic>
#include <condition_variable>
#include <mutex>

#include "pipeline.cuh"

namespace rg {

void stream_copy(void* dst, const void* src, size_t n);  // host_copy.cpp: non-temporal stores

// Persistent worker threads for the staging memcpy of the pageable host paths.  (Spawning and joining threads
// per call cost ~0.3 ms: more than the copy of an 8 MiB piece.)  The calling thread takes part in the copy; the
// pool grows on demand and lives until the process exits.  One copy at a time (calls are serialised).
class CopyPool {
 public:
  void run(unsigned char* dst, const unsigned char* src, size_t bytes, int threads) {
    std::lock_guard<std::mutex> serial(call_mutex_);
    grow(threads - 1);
    // 64 KiB-aligned parts, four per thread: late starters take fewer parts
    const size_t parts = static_cast<size_t>(threads) * 4;
    const size_t part = ((bytes + parts - 1) / parts + 65535) / 65536 * 65536;
    {
      std::lock_guard<std::mutex> lk(m_);
      dst_ = dst;
      src_ = src;
      bytes_ = bytes;
      part_ = part;
      n_parts_ = (bytes + part - 1) / part;
      next_.store(0, std::memory_order_relaxed);
      helpers_ = threads - 1;
      active_ = helpers_;
      ++generation_;
    }
    cv_work_.notify_all();
    work();
    std::unique_lock<std::mutex> lk(m_);
    cv_done_.wait(lk, [&] { return active_ == 0; });
  }

  ~CopyPool() {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
    }
    cv_work_.notify_all();
    for (auto& t : workers_) t.join();
  }

 private:
  void work() {
    for (;;) {
      const size_t i = next_.fetch_add(1, std::memory_order_relaxed);
      if (i >= n_parts_) break;
      const size_t off = i * part_;
      const size_t n = bytes_ - off < part_ ? bytes_ - off : part_;
      stream_copy(dst_ + off, src_ + off, n);
    }
  }
  void grow(int helpers) {
    while (static_cast<int>(workers_.size()) < helpers) {
      const int id = static_cast<int>(workers_.size());
      workers_.emplace_back([this, id] {
        unsigned long long seen = 0;
        for (;;) {
          {
            std::unique_lock<std::mutex> lk(m_);
            cv_work_.wait(lk, [&] { return stop_ || (generation_ != seen && id < helpers_); });
            if (stop_) return;
            seen = generation_;
          }
          work();
          std::lock_guard<std::mutex> lk(m_);
          if (--active_ == 0) cv_done_.notify_one();
        }
      });
    }
  }

  std::mutex call_mutex_, m_;
  std::condition_variable cv_work_, cv_done_;
  std::vector<std::thread> workers_;
  unsigned char* dst_ = nullptr;
  const unsigned char* src_ = nullptr;
  size_t bytes_ = 0, part_ = 0, n_parts_ = 0;
  std::atomic<size_t> next_{0};
  int helpers_ = 0, active_ = 0;
  unsigned long long generation_ = 0;
  bool stop_ = false;
};

// The pool is intentionally leaked (its threads must not be joined at exit), and a fork()ed child -- which
// inherits the object but none of its threads -- starts a fresh one.
static std::atomic<CopyPool*> g_copy_pool{nullptr};
static std::mutex g_copy_pool_mutex;

inline CopyPool& copy_pool() {
  CopyPool* p = g_copy_pool.load(std::memory_order_acquire);
  if (!p) {
    std::lock_guard<std::mutex> lk(g_copy_pool_mutex);
    p = g_copy_pool.load(std::memory_order_acquire);
    if (!p) {
      static bool hooked = false;
      if (!hooked) {
        pthread_atfork(nullptr, nullptr, [] {
          g_copy_pool.store(nullptr, std::memory_order_release);
          new (&g_copy_pool_mutex) std::mutex();  // the parent may have held it at fork time
        });
        hooked = true;
      }
      p = new CopyPool();
      g_copy_pool.store(p, std::memory_order_release);
    }
  }
  return *p;
}

}  // namespace rg

extern "C" {

int rg_pipeline_create(rg_pipeline_t** out, int32_t n_slots) {
  if (!out) return RG_ERR_NULL;
  *out = nullptr;
  rg_pipeline* p = new (std::nothrow) rg_pipeline();
  if (!p) return static_cast<int>(cudaErrorMemoryAllocation);
  const int rc = p->create(n_slots);
  if (rc != RG_OK) {
    p->destroy();
    delete p;
    return rc;
  }
  *out = p;
  return RG_OK;
}

int rg_pipeline_destroy(rg_pipeline_t* p) {
  if (!p) return RG_OK;
  p->destroy();
  delete p;
  return RG_OK;
}

int32_t rg_pipeline_slots(const rg_pipeline_t* p) { return p ? p->n_slots : 0; }

void* rg_pipeline_stream(rg_pipeline_t* p, int32_t which) {
  if (!p) return nullptr;
  return which == 0 ? p->s_in : which == 1 ? p->s_k : which == 2 ? p->s_out : nullptr;
}

int rg_pipeline_h2d(rg_pipeline_t* p, int32_t slot, void* dst_device, const void* src_host, size_t bytes) {
  if (!p || (bytes && (!dst_device || !src_host))) return RG_ERR_NULL;
  return p->h2d(slot, dst_device, src_host, bytes);
}

int rg_pipeline_compute_begin(rg_pipeline_t* p, int32_t slot) { return p ? p->compute_begin(slot) : RG_ERR_NULL; }

int rg_pipeline_compute_end(rg_pipeline_t* p, int32_t slot) { return p ? p->compute_end(slot) : RG_ERR_NULL; }

int rg_pipeline_d2h(rg_pipeline_t* p, int32_t slot, void* dst_host, const void* src_device, size_t bytes) {
  if (!p || (bytes && (!dst_host || !src_device))) return RG_ERR_NULL;
  return p->d2h(slot, dst_host, src_device, bytes);
}

int rg_pipeline_h2d_pageable(rg_pipeline_t* p, int32_t slot, void* dst_device, const void* src_host, size_t bytes,
                             void* stage_pinned, size_t stage_bytes, int32_t n_threads) {
  if (!p || (bytes && (!dst_device || !src_host || !stage_pinned))) return RG_ERR_NULL;
  return p->h2d_pageable(slot, dst_device, src_host, bytes, stage_pinned, stage_bytes, n_threads);
}

int rg_pipeline_d2h_pageable(rg_pipeline_t* p, int32_t slot, void* dst_host, const void* src_device, size_t bytes,
                             void* stage_pinned, size_t stage_bytes, int32_t n_threads) {
  if (!p || (bytes && (!dst_host || !src_device || !stage_pinned))) return RG_ERR_NULL;
  return p->d2h_pageable(slot, dst_host, src_device, bytes, stage_pinned, stage_bytes, n_threads);
}

int rg_pipeline_wait_stream(rg_pipeline_t* p, void* stream) {
  return p ? p->wait_stream(static_cast<cudaStream_t>(stream)) : RG_ERR_NULL;
}

int rg_pipeline_sync(rg_pipeline_t* p, int32_t stage, int32_t slot) {
  if (!p) return RG_ERR_NULL;
  if (slot < 0 || slot >= p->n_slots || stage < 0 || stage > 2) return RG_ERR_SHAPE;
  cudaEvent_t e = stage == 0 ? p->in_done[slot] : stage == 1 ? p->k_done[slot] : p->out_done[slot];
  return static_cast<int>(cudaEventSynchronize(e));
}

int rg_pipeline_drain(rg_pipeline_t* p) { return p ? p->drain() : RG_ERR_NULL; }

int rg_host_copy(void* dst, const void* src, size_t bytes, int32_t n_threads) {
  if (bytes == 0) return RG_OK;
  if (!dst || !src) return RG_ERR_NULL;
  constexpr size_t kMinPerThread = size_t(4) << 20;
  size_t want = n_threads > 0 ? static_cast<size_t>(n_threads) : 1;
  if (want > 64) want = 64;
  if (want > (bytes + kMinPerThread - 1) / kMinPerThread) want = (bytes + kMinPerThread - 1) / kMinPerThread;
  if (want <= 1) {
    rg::stream_copy(dst, src, bytes);
    return RG_OK;
  }
  rg::copy_pool().run(static_cast<unsigned char*>(dst), static_cast<const unsigned char*>(src), bytes,
                      static_cast<int>(want));
  return RG_OK;
}

}  // extern "C"
