// C ABI of the persistent streaming pipeline + the threaded host copy used to stage pageable memory.
#include <cstring>
#include <new>
#include <thread>
#include <vector>

#include "pipeline.cuh"

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
    std::memcpy(dst, src, bytes);
    return RG_OK;
  }
  // 4 KiB aligned shares: whole pages per thread
  const size_t share = ((bytes + want - 1) / want + 4095) / 4096 * 4096;
  std::vector<std::thread> pool;
  pool.reserve(want - 1);
  unsigned char* d = static_cast<unsigned char*>(dst);
  const unsigned char* s = static_cast<const unsigned char*>(src);
  for (size_t t = 1; t < want; ++t) {
    const size_t a = t * share;
    if (a >= bytes) break;
    const size_t n = (bytes - a < share) ? bytes - a : share;
    pool.emplace_back([=] { std::memcpy(d + a, s + a, n); });
  }
  std::memcpy(d, s, share < bytes ? share : bytes);
  for (auto& th : pool) th.join();
  return RG_OK;
}

}  // extern "C"
