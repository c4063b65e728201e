// TMA row-ring conservative kernel (sm_100a): persistent, warp-specialised.
//
//   producer (1 elected thread of warp 8)   streams whole source rows HBM -> shared-memory ring
//                                           with cp.async.bulk (TMA, SASS: UBLKCP) and arms one
//                                           "full" mbarrier per ring slot with the row's byte count
//   consumers (warps 0-7)                   wait on "full", contract the band's rows along latitude
//                                           into double-buffered t / v column arrays (phase 1),
//                                           hand dead rows back through "empty" mbarriers, then
//                                           contract t / v along longitude, normalise, mask and
//                                           store one output row (phase 2)
//
// A unit of work is (slab b, run of consecutive target rows); the ring keeps rows alive across
// the target rows that share them, so every source row crosses HBM->SM exactly once per run and
// the producer prefetches across row, run and slab boundaries (no load/compute phase bubbles).
// The load order / liveness schedule is precomputed on the host (rg_row_ring_t).
#pragma once

#include "common.cuh"

namespace rg {

struct RingDev {
  const int32_t* __restrict__ sched_row;
  const int32_t* __restrict__ tap_pos;
  const int32_t* __restrict__ need;
  const int32_t* __restrict__ release;
  const int32_t* __restrict__ run_first;
  const int32_t* __restrict__ run_sched;
  int n_runs, n_sched, max_live, pad_shift;
};

constexpr int kRingConsumers = 256;
constexpr int kRingThreads = kRingConsumers + 32;
constexpr int kRingMaxSlots = 16;
constexpr int kTapChunk = 8;

// Shared-memory carve-up, computed identically on host and device.
struct RingLayout {
  int ring_slots, row_stride, w_slots, tv_stride;
  size_t off_bar, off_tv, off_cw, off_cidx, off_ccnt, total;
};

template <typename T>
__host__ __device__ inline RingLayout ring_layout(int W, int Wo, int Kw, int pad_shift, bool skipna,
                                                  int ring_slots) {
  RingLayout L;
  L.ring_slots = ring_slots;
  L.row_stride = (W * static_cast<int>(sizeof(T)) + 127) / 128 * 128;
  L.w_slots = ((W + (pad_shift < 31 ? (W >> pad_shift) : 0) + 4) + 3) / 4 * 4;
  L.tv_stride = (skipna ? 2 : 1) * L.w_slots;
  size_t off = static_cast<size_t>(ring_slots) * L.row_stride;
  L.off_bar = off;
  off += static_cast<size_t>(2 * kRingMaxSlots) * sizeof(unsigned long long);
  L.off_tv = off;
  off += static_cast<size_t>(2) * L.tv_stride * sizeof(T);
  L.off_cw = off;
  off += (static_cast<size_t>(Kw) * Wo * sizeof(T) + 15) / 16 * 16;
  L.off_cidx = off;
  off += (static_cast<size_t>(Kw) * Wo * sizeof(int32_t) + 15) / 16 * 16;
  L.off_ccnt = off;
  off += (static_cast<size_t>(Wo) * sizeof(int32_t) + 15) / 16 * 16;
  L.total = off;
  return L;
}

// ------------------------------------------------------------------ mbarrier / TMA PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// 1-D bulk copy global -> shared, completion signalled on an mbarrier (TMA, no tensor map needed).
__device__ __forceinline__ void tma_load_row(void* dst, const void* src, uint32_t bytes,
                                             unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void consumer_barrier() {
  asm volatile("bar.sync 1, %0;" ::"n"(kRingConsumers) : "memory");
}

template <typename T, bool SKIPNA>
__global__ void __launch_bounds__(kRingThreads, 1)
conservative_ring_kernel(const T* __restrict__ x, T* __restrict__ y, long long batch, long long xbs,
                         long long xrs, long long ybs, Band<T> rows, Band<T> cols, RingDev ring, T thr,
                         int ring_slots) {
  constexpr int VEC = 16 / sizeof(T);
  extern __shared__ __align__(128) unsigned char smem[];
  const int W = cols.n_src, Wo = cols.n_out, Kw = cols.k_max, Ho = rows.n_out;
  const int pad_shift = ring.pad_shift;
  const RingLayout L = ring_layout<T>(W, Wo, Kw, pad_shift, SKIPNA, ring_slots);
  unsigned char* s_ring = smem;
  unsigned long long* s_full = reinterpret_cast<unsigned long long*>(smem + L.off_bar);
  unsigned long long* s_empty = s_full + kRingMaxSlots;
  T* s_tv = reinterpret_cast<T*>(smem + L.off_tv);
  T* s_cw = reinterpret_cast<T*>(smem + L.off_cw);
  int32_t* s_cidx = reinterpret_cast<int32_t*>(smem + L.off_cidx);
  int32_t* s_ccnt = reinterpret_cast<int32_t*>(smem + L.off_ccnt);

  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < ring_slots; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], 1);
    }
    fence_mbar_init();
  }
  // longitude band -> shared memory once per (persistent) CTA; indices pre-mapped to padded slots
  for (int i = tid; i < Kw * Wo; i += kRingThreads) {
    const int j = cols.idx[i];
    s_cidx[i] = j + (pad_shift < 31 ? (j >> pad_shift) : 0);
    s_cw[i] = cols.w[i];
  }
  for (int i = tid; i < Wo; i += kRingThreads) s_ccnt[i] = cols.cover[i] ? cols.cnt[i] : -1;
  __syncthreads();

  const long long tasks = batch * ring.n_runs;
  const uint32_t row_bytes = static_cast<uint32_t>(W) * sizeof(T);

  // ================================================================ producer warp
  if (tid >= kRingConsumers) {
    if (tid == kRingConsumers) {
      int slot = 0;
      uint32_t parity = 1;  // waiting on the "previous" phase of a fresh barrier returns at once
      for (long long task = blockIdx.x; task < tasks; task += gridDim.x) {
        const int run = static_cast<int>(task % ring.n_runs);
        const T* xb = x + (task / ring.n_runs) * xbs;
        const int p1 = ring.run_sched[run + 1];
        for (int p = ring.run_sched[run]; p < p1; ++p) {
          mbar_wait(&s_empty[slot], parity);
          mbar_arrive_expect_tx(&s_full[slot], row_bytes);
          tma_load_row(s_ring + static_cast<size_t>(slot) * L.row_stride,
                       xb + static_cast<long long>(ring.sched_row[p]) * xrs, row_bytes, &s_full[slot]);
          if (++slot == ring_slots) { slot = 0; parity ^= 1; }
        }
      }
    }
    return;
  }

  // ================================================================ consumer warps
  const T nan = quiet_nan<T>();
  uint32_t gbase = 0;                       // ring position of the current run's first row
  uint32_t landed = 0, landed_slot = 0, landed_parity = 0;
  uint32_t released = 0, released_slot = 0;
  int buf = 0;

  for (long long task = blockIdx.x; task < tasks; task += gridDim.x) {
    const int run = static_cast<int>(task % ring.n_runs);
    const long long b = task / ring.n_runs;
    const int k1 = ring.run_first[run + 1];
    for (int k = ring.run_first[run]; k < k1; ++k) {
      const bool covered = rows.cover[k] != 0;
      const int cnt = covered ? rows.cnt[k] : 0;
      T* s_t = s_tv + buf * L.tv_stride;
      T* s_v = s_t + L.w_slots;

      // ---- wait until every row this output needs has landed in the ring
      const uint32_t need_g = gbase + static_cast<uint32_t>(ring.need[k]);
      while (landed < need_g) {
        mbar_wait(&s_full[landed_slot], landed_parity);
        ++landed;
        if (++landed_slot == static_cast<uint32_t>(ring_slots)) { landed_slot = 0; landed_parity ^= 1; }
      }

      // ---- phase 1: latitude contraction out of the ring
      for (int c0 = 0; c0 < cnt || c0 == 0; c0 += kTapChunk) {
        const unsigned char* rowp[kTapChunk];
        T w[kTapChunk];
#pragma unroll
        for (int i = 0; i < kTapChunk; ++i) {
          if (c0 + i < cnt) {
            const uint32_t pos = gbase + static_cast<uint32_t>(ring.tap_pos[(c0 + i) * Ho + k]);
            rowp[i] = s_ring + static_cast<size_t>(pos % static_cast<uint32_t>(ring_slots)) * L.row_stride;
            w[i] = rows.w[(c0 + i) * Ho + k];
          }
        }
        for (int j = tid * VEC; j < W; j += kRingConsumers * VEC) {
          T acc[VEC], vac[VEC];
#pragma unroll
          for (int e = 0; e < VEC; ++e) { acc[e] = T(0); vac[e] = T(0); }
          Pack<T, VEC> val[kTapChunk];
#pragma unroll
          for (int i = 0; i < kTapChunk; ++i) {
            if (c0 + i < cnt) val[i] = *reinterpret_cast<const Pack<T, VEC>*>(rowp[i] + j * sizeof(T));
          }
#pragma unroll
          for (int i = 0; i < kTapChunk; ++i) {
            if (c0 + i < cnt) {
#pragma unroll
              for (int e = 0; e < VEC; ++e) {
                const T xv = val[i].v[e];
                if (SKIPNA) {
                  const bool ok = (xv == xv);
                  acc[e] = fma(w[i], ok ? xv : T(0), acc[e]);
                  vac[e] += ok ? w[i] : T(0);
                } else {
                  acc[e] = fma(w[i], xv, acc[e]);
                }
              }
            }
          }
#pragma unroll
          for (int e = 0; e < VEC; ++e) {
            if (j + e < W) {
              const int s = (j + e) + (pad_shift < 31 ? ((j + e) >> pad_shift) : 0);
              if (c0 == 0) {
                s_t[s] = acc[e];
                if (SKIPNA) s_v[s] = vac[e];
              } else {
                s_t[s] += acc[e];
                if (SKIPNA) s_v[s] += vac[e];
              }
            }
          }
        }
      }
      consumer_barrier();  // t / v complete, ring reads of this output complete

      // ---- hand dead rows back to the producer
      const uint32_t rel_g = gbase + static_cast<uint32_t>(ring.release[k]);
      while (released < rel_g) {
        if (tid == 0) mbar_arrive(&s_empty[released_slot]);
        ++released;
        if (++released_slot == static_cast<uint32_t>(ring_slots)) released_slot = 0;
      }

      // ---- phase 2: longitude contraction, normalise, mask, store
      T* yrow = y + b * ybs + static_cast<long long>(k) * Wo;
      for (int l = tid; l < Wo; l += kRingConsumers) {
        const int c = s_ccnt[l];
        T out = nan;
        if (covered && c >= 0) {
          T so = T(0), sv = T(0);
          for (int t0 = 0; t0 < c; t0 += 4) {
            T tw[4], tt[4], tvv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (t0 + i < c) {
                const int s = s_cidx[(t0 + i) * Wo + l];
                tw[i] = s_cw[(t0 + i) * Wo + l];
                tt[i] = s_t[s];
                if (SKIPNA) tvv[i] = s_v[s];
              }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (t0 + i < c) {
                so = fma(tw[i], tt[i], so);
                if (SKIPNA) sv = fma(tw[i], tvv[i], sv);
              }
            }
          }
          if (SKIPNA) {
            out = (sv >= thr) ? so / sv : nan;
          } else {
            out = so;
          }
        }
        st_stream(yrow + l, out);
      }
      buf ^= 1;  // double-buffered t / v: the next phase 1 may start while others finish phase 2
    }
    gbase += static_cast<uint32_t>(ring.run_sched[run + 1] - ring.run_sched[run]);
  }
}

}  // namespace rg
