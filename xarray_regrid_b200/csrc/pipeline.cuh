// Persistent host <-> device streaming pipeline (rg_pipeline_t of include/regrid_b200.h).
//
// Three non-blocking streams (H2D, compute, D2H) and one event triple per slot, created ONCE and reused
// for any number of chunks / calls.  A slot is a caller-owned set of device buffers (input chunk, output
// chunk, scratch); the pipeline only orders the work:
//
//   h2d(slot)            s_in  waits k_done[slot]   (the kernel that read the slot's previous input)
//                        copy, record in_done[slot]
//   compute_begin(slot)  s_k   waits in_done[slot] and out_done[slot] (previous result has left the slot)
//   compute_end(slot)    record k_done[slot] on s_k
//   d2h(slot)            s_out waits k_done[slot]; copy, record out_done[slot]
//
// Stands in for the reference's dask chunk scheduling (methods/conservative.py:263-302, interp.py:43).
#pragma once

#include <cuda_runtime.h>

#include "common.cuh"

constexpr int kPipelineMaxSlots = 8;
constexpr int kPipelinePieces = 4;  // page-locked staging pieces in flight per direction (pageable paths)

extern "C" int rg_host_copy(void* dst, const void* src, size_t bytes, int32_t n_threads);

struct rg_pipeline {
  int n_slots = 0;
  cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
  cudaEvent_t in_done[kPipelineMaxSlots] = {}, k_done[kPipelineMaxSlots] = {}, out_done[kPipelineMaxSlots] = {};
  cudaEvent_t ext = nullptr;  // dependency on a caller stream (rg_pipeline_wait_stream)
  cudaEvent_t piece_in[kPipelinePieces] = {}, piece_out[kPipelinePieces] = {};  // staging pieces (pageable paths)

  int create(int slots) {
    if (slots < 1 || slots > kPipelineMaxSlots) return RG_ERR_SHAPE;
    n_slots = slots;
    RG_CUDA(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
    RG_CUDA(cudaStreamCreateWithFlags(&s_k, cudaStreamNonBlocking));
    RG_CUDA(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
    for (int i = 0; i < slots; ++i) {
      RG_CUDA(cudaEventCreateWithFlags(&in_done[i], cudaEventDisableTiming));
      RG_CUDA(cudaEventCreateWithFlags(&k_done[i], cudaEventDisableTiming));
      RG_CUDA(cudaEventCreateWithFlags(&out_done[i], cudaEventDisableTiming));
    }
    RG_CUDA(cudaEventCreateWithFlags(&ext, cudaEventDisableTiming));
    for (int i = 0; i < kPipelinePieces; ++i) {
      RG_CUDA(cudaEventCreateWithFlags(&piece_in[i], cudaEventDisableTiming));
      RG_CUDA(cudaEventCreateWithFlags(&piece_out[i], cudaEventDisableTiming));
    }
    return RG_OK;
  }

  // Host-blocks until everything queued on the three streams has finished.  Always drains all three, even
  // after an error, so that no copy is in flight when the caller frees or reuses its buffers.
  int drain() {
    cudaError_t e0 = s_in ? cudaStreamSynchronize(s_in) : cudaSuccess;
    cudaError_t e1 = s_k ? cudaStreamSynchronize(s_k) : cudaSuccess;
    cudaError_t e2 = s_out ? cudaStreamSynchronize(s_out) : cudaSuccess;
    return static_cast<int>(e0 != cudaSuccess ? e0 : (e1 != cudaSuccess ? e1 : e2));
  }

  void destroy() {
    drain();
    for (int i = 0; i < kPipelineMaxSlots; ++i) {
      if (in_done[i]) cudaEventDestroy(in_done[i]);
      if (k_done[i]) cudaEventDestroy(k_done[i]);
      if (out_done[i]) cudaEventDestroy(out_done[i]);
      in_done[i] = k_done[i] = out_done[i] = nullptr;
    }
    if (ext) cudaEventDestroy(ext);
    for (int i = 0; i < kPipelinePieces; ++i) {
      if (piece_in[i]) cudaEventDestroy(piece_in[i]);
      if (piece_out[i]) cudaEventDestroy(piece_out[i]);
      piece_in[i] = piece_out[i] = nullptr;
    }
    if (s_in) cudaStreamDestroy(s_in);
    if (s_k) cudaStreamDestroy(s_k);
    if (s_out) cudaStreamDestroy(s_out);
    ext = nullptr;
    s_in = s_k = s_out = nullptr;
  }

  int h2d(int slot, void* dst, const void* src, size_t bytes) {
    if (slot < 0 || slot >= n_slots) return RG_ERR_SHAPE;
    RG_CUDA(cudaStreamWaitEvent(s_in, k_done[slot], 0));
    if (bytes) RG_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s_in));
    RG_CUDA(cudaEventRecord(in_done[slot], s_in));
    return RG_OK;
  }
  int compute_begin(int slot) {
    if (slot < 0 || slot >= n_slots) return RG_ERR_SHAPE;
    RG_CUDA(cudaStreamWaitEvent(s_k, in_done[slot], 0));
    RG_CUDA(cudaStreamWaitEvent(s_k, out_done[slot], 0));
    return RG_OK;
  }
  int compute_end(int slot) {
    if (slot < 0 || slot >= n_slots) return RG_ERR_SHAPE;
    RG_CUDA(cudaEventRecord(k_done[slot], s_k));
    return RG_OK;
  }
  int d2h(int slot, void* dst, const void* src, size_t bytes) {
    if (slot < 0 || slot >= n_slots) return RG_ERR_SHAPE;
    RG_CUDA(cudaStreamWaitEvent(s_out, k_done[slot], 0));
    if (bytes) RG_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s_out));
    RG_CUDA(cudaEventRecord(out_done[slot], s_out));
    return RG_OK;
  }
  // Pageable source: the chunk is cut into pieces that are copied (multi-threaded memcpy) into a ring of
  // kPipelinePieces page-locked staging pieces and sent with one cudaMemcpyAsync each; the host only blocks
  // when it laps the ring, so the staging copy of piece k + 1 overlaps the DMA of piece k.
  int h2d_pageable(int slot, void* dst, const void* src, size_t bytes, void* stage, size_t stage_bytes,
                   int n_threads) {
    if (slot < 0 || slot >= n_slots) return RG_ERR_SHAPE;
    const size_t piece = stage_bytes / kPipelinePieces / 4096 * 4096;
    if (bytes && piece == 0) return RG_ERR_SHAPE;
    RG_CUDA(cudaStreamWaitEvent(s_in, k_done[slot], 0));
    unsigned char* d = static_cast<unsigned char*>(dst);
    const unsigned char* s = static_cast<const unsigned char*>(src);
    unsigned char* st = static_cast<unsigned char*>(stage);
    int k = 0;
    for (size_t off = 0; off < bytes; off += piece, k = (k + 1) % kPipelinePieces) {
      const size_t n = bytes - off < piece ? bytes - off : piece;
      RG_CUDA(cudaEventSynchronize(piece_in[k]));  // the piece's previous DMA has finished (no-op if unused)
      rg_host_copy(st + k * piece, s + off, n, n_threads);
      RG_CUDA(cudaMemcpyAsync(d + off, st + k * piece, n, cudaMemcpyHostToDevice, s_in));
      RG_CUDA(cudaEventRecord(piece_in[k], s_in));
    }
    RG_CUDA(cudaEventRecord(in_done[slot], s_in));
    return RG_OK;
  }
  // Pageable result: D2H in pieces through the staging ring, copied out on the host as they land.  Blocks the
  // host until the whole chunk is in `dst` (queue the next chunk's H2D and kernels BEFORE calling this).
  int d2h_pageable(int slot, void* dst, const void* src, size_t bytes, void* stage, size_t stage_bytes,
                   int n_threads) {
    if (slot < 0 || slot >= n_slots) return RG_ERR_SHAPE;
    const size_t piece = stage_bytes / kPipelinePieces / 4096 * 4096;
    if (bytes && piece == 0) return RG_ERR_SHAPE;
    RG_CUDA(cudaStreamWaitEvent(s_out, k_done[slot], 0));
    unsigned char* d = static_cast<unsigned char*>(dst);
    const unsigned char* s = static_cast<const unsigned char*>(src);
    unsigned char* st = static_cast<unsigned char*>(stage);
    const size_t n_pieces = (bytes + piece - 1) / (piece ? piece : 1);
    size_t issued = 0, done = 0;
    while (done < n_pieces) {
      while (issued < n_pieces && issued - done < static_cast<size_t>(kPipelinePieces)) {
        const size_t off = issued * piece;
        const size_t n = bytes - off < piece ? bytes - off : piece;
        const int k = static_cast<int>(issued % kPipelinePieces);
        RG_CUDA(cudaMemcpyAsync(st + k * piece, s + off, n, cudaMemcpyDeviceToHost, s_out));
        RG_CUDA(cudaEventRecord(piece_out[k], s_out));
        ++issued;
        if (issued == n_pieces) RG_CUDA(cudaEventRecord(out_done[slot], s_out));
      }
      const size_t off = done * piece;
      const size_t n = bytes - off < piece ? bytes - off : piece;
      const int k = static_cast<int>(done % kPipelinePieces);
      RG_CUDA(cudaEventSynchronize(piece_out[k]));
      rg_host_copy(d + off, st + k * piece, n, n_threads);
      ++done;
    }
    if (n_pieces == 0) RG_CUDA(cudaEventRecord(out_done[slot], s_out));
    return RG_OK;
  }
  // every pipeline stream waits for the work queued so far on `stream` (the caller's): buffers the caller
  // produced or recycled on its own stream are safe to use from the pipeline afterwards
  int wait_stream(cudaStream_t stream) {
    RG_CUDA(cudaEventRecord(ext, stream));
    RG_CUDA(cudaStreamWaitEvent(s_in, ext, 0));
    RG_CUDA(cudaStreamWaitEvent(s_k, ext, 0));
    RG_CUDA(cudaStreamWaitEvent(s_out, ext, 0));
    return RG_OK;
  }
};
