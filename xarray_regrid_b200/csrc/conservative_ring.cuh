// TMA row-ring conservative kernel (sm_100a): persistent, warp-specialised, barrier-free.
//
//   producer (one elected thread of warp n_warps)
//       streams whole source rows HBM -> shared-memory ring with cp.async.bulk (TMA, SASS: UBLKCP)
//       and arms one "full" mbarrier per ring slot with the row's byte count; a slot is reused once
//       every consumer warp has published (st.release) that it is done with the row it held
//   sync (one elected thread of warp n_warps + 1)
//       waits on the "full" mbarriers in order and publishes the number of landed rows in one
//       shared counter, so a consumer's wait is a single ld.acquire instead of per-row barrier math
//   consumer warps (all other warps), each owning strips of <= 32 target columns
//       poll the landed counter; contract the band's rows along latitude for the source columns their strip
//       needs (its taps + halo) into a warp-private shared buffer of {t, v} pairs (phase 1);
//       __syncwarp; contract along longitude, normalise, mask and store their part of the output
//       row (phase 2); publish their progress so the producer can recycle dead rows
//
// There is no CTA-wide barrier after start-up: warps run decoupled, limited only by the ring.
// A unit of work is (slab b, run of consecutive target rows); the ring keeps rows alive across the
// target rows that share them, so every source row crosses HBM->SM exactly once per run and the
// producer prefetches across row, run and slab boundaries.  Load order / liveness (rows) and the
// strip decomposition (columns) are precomputed on the host (rg_row_ring_t).
#pragma once

#include "common.cuh"

namespace rg {

struct RingDev {
  const int32_t* __restrict__ sched_row;
  const int4* __restrict__ krec;   // per output row: {need, release, kcnt, 8 x 4-bit tap_back}
  const void* __restrict__ wrec;   // per output row: 8 tap weights (zero padded), output-major
  const int32_t* __restrict__ run_first;
  const int32_t* __restrict__ run_sched;
  const int32_t* __restrict__ strip_l0;
  const int32_t* __restrict__ strip_c0;
  const int32_t* __restrict__ strip_nc;
  int n_runs, n_sched, max_live, pad_shift, n_strips, strip_cols_max, n_warps, lane_cols;
  const void* __restrict__ pole;  // [batch, 2, W] synthetic pole rows (scheduled as rows n_src_rows, + 1) or NULL
  int n_src_rows;
  int zero_fill;  // !SKIPNA kernels: a NaN tap counts as 0 (apply_weights with skipna=False) instead of poisoning
  // column panels: a task is (slab, run, panel); the ring rows hold the panel's source columns only
  const int32_t* __restrict__ panel_strip0;
  const int32_t* __restrict__ panel_c0;
  const int32_t* __restrict__ panel_nc;
  int n_panels, panel_cols_max;
  // block layout of the register-resident kernel (lane_cols == 2)
  const int32_t* __restrict__ strip_b0;
  const int32_t* __restrict__ block_l0;
  int halo_l, halo_r;
  int lane_split;  // lanes kernel, one target per lane: split points of the latitude pass (0 - 2, see SPLIT)
  int l2_hint;     // producer: bulk copies carry an L2 evict-first hint
};

constexpr int kRingMaxSlots = 32;  // ring depth cap (a tap reaches at most 15 rows back: 4-bit tap_back)
constexpr int kRingMaxLive = 15;
constexpr int kRingMaxTaps = 8;   // row taps per output handled by the ring kernel
constexpr int kRingMaxWarps = 24; // consumer warps
constexpr int kLaneWrecSmall = 6; // staged weights per target row in the lanes instances of <= 5 row taps
template <int RMAX> struct LaneSched { using type = int32_t; };
template <> struct LaneSched<5> { using type = uint16_t; };  // (host contract: source rows + 2 pole rows < 65536)

// Shared-memory carve-up, computed identically on host and device.
struct RingLayout {
  int ring_slots, row_stride, strip_slots;
  size_t strip_bytes;  // one warp-private {t, v} buffer
  size_t off_bar, off_flags, off_strip, total;
};

template <typename T>
__host__ __device__ inline RingLayout ring_layout(int W, int pad_shift, bool skipna, int strip_cols_max,
                                                  int n_warps, int ring_slots) {
  RingLayout L;
  L.ring_slots = ring_slots;
  L.row_stride = (W * static_cast<int>(sizeof(T)) + 127) / 128 * 128;
  L.strip_slots = strip_cols_max + (pad_shift < 31 ? (strip_cols_max >> pad_shift) : 0) + 1;
  L.strip_bytes = (static_cast<size_t>(L.strip_slots) * (skipna ? 2 : 1) * sizeof(T) + 127) / 128 * 128;
  size_t off = static_cast<size_t>(ring_slots) * L.row_stride;
  L.off_bar = off;
  off += static_cast<size_t>(kRingMaxSlots) * sizeof(unsigned long long);
  L.off_flags = off;  // [0] = rows landed, [1 + w] = rows released by consumer warp w
  off += 128;
  L.off_strip = off;
  off += static_cast<size_t>(n_warps) * L.strip_bytes;
  L.total = off;
  return L;
}

// ------------------------------------------------------------------ mbarrier / TMA / flag PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar_addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok)
      : "r"(bar_addr), "r"(parity)
      : "memory");
  return ok != 0;
}
// 1-D bulk copy global -> shared, completion signalled on an mbarrier (TMA, no tensor map needed).
__device__ __forceinline__ void tma_load_row(uint32_t dst_addr, const void* src, uint32_t bytes,
                                             uint32_t bar_addr) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_addr),
      "l"(src), "r"(bytes), "r"(bar_addr)
      : "memory");
}
// the same with an L2 eviction-priority hint for the source lines (streamed once, never reused)
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_load_row_hint(uint32_t dst_addr, const void* src, uint32_t bytes,
                                                  uint32_t bar_addr, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          dst_addr),
      "l"(src), "r"(bytes), "r"(bar_addr), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_smem(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_relaxed_smem(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.relaxed.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acq_rel_cta() { asm volatile("fence.acq_rel.cta;" ::: "memory"); }
// Progress flag of a consumer warp whose ring reads have already been CONSUMED (their values fed instructions
// that precede this store in program order, so the loads have completed): a relaxed store is enough and, unlike
// st.release (MEMBAR.ALL.CTA + STS), it does not wait for the warp's output stores still in flight.  The
// producer turns its poll into an acquire with one fence before it lets the TMA overwrite a slot.
__device__ __forceinline__ void st_relaxed_smem(uint32_t addr, uint32_t v) {
  asm volatile("st.relaxed.cta.shared::cta.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_smem(uint32_t addr, uint32_t v) {
  asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// so / sv for the normalisation step.  fp64: hardware reciprocal seed + two Newton steps + one residual
// correction of the quotient (9 instructions instead of the ~35 of the IEEE division routine).  sv is a
// valid mass in [valid_thr, 1 + eps] and so a weighted mean of the data, so the special-case handling of
// the full routine is not needed; the result is within 1 ulp of the correctly rounded quotient (and equal
// to it in all but rare cases), far inside the 1e-12 parity tolerance.
__device__ __forceinline__ double normalise_div(double so, double sv) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(sv));
  double e = fma(-sv, r, 1.0);
  r = fma(r, e, r);
  e = fma(-sv, r, 1.0);
  r = fma(r, e, r);
  const double q = so * r;
  return fma(fma(-sv, q, so), r, q);
}
__device__ __forceinline__ float normalise_div(float so, float sv) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(sv));  // 1 ulp seed
  r = fmaf(r, fmaf(-sv, r, 1.0f), r);
  const float q = so * r;
  return fmaf(fmaf(-sv, q, so), r, q);  // residual correction: within 1 ulp of so / sv, no subroutine call
}

// {t, v} entry of the warp-private strip buffer.
template <typename T, bool SKIPNA> struct TV;
template <typename T> struct alignas(2 * sizeof(T)) TV<T, true> { T t, v; };
template <typename T> struct TV<T, false> { T t; };

// acc += w * x and vac += w, both skipped when x is NaN: the reference's fillna(0) / notnull() contributions
// (w * 0 adds nothing).  fp64: five instructions per element -- one NaN test, one 32-bit select per operand
// (ptxas turns predicated DFMA / DADD, and the C++ `ok ? x : 0`, into a full-width select pair: 7) and the two
// accumulates.  Only the HIGH word of a skipped value is cleared: a canonical NaN (low word 0) becomes exactly
// 0.0; a NaN carrying payload bits in its low word becomes a denormal below 2^-1042 that can never reach the
// rounding of a result.
__device__ __forceinline__ void accum_valid(double& acc, double& vac, double w, double x, double& one) {
  // valid mass: vac += w * (ok ? 1.0 : 0.0).  `one` is a caller-held scratch whose LOW word stays 0: only its
  // high word is rewritten (one 32-bit select), so the 0 / 1 factor costs no register moves (w * 1.0 is exact)
  asm("{\n .reg .pred p;\n .reg .b32 lo, hi, ol, oh;\n .reg .f64 xc;\n"
      " setp.num.f64 p, %3, %3;\n"
      " mov.b64 {lo, hi}, %3;\n selp.b32 hi, hi, 0, p;\n mov.b64 xc, {lo, hi};\n"
      " mov.b64 {ol, oh}, %2;\n selp.b32 oh, 0x3ff00000, 0, p;\n mov.b64 %2, {ol, oh};\n"
      " fma.rn.f64 %0, %4, xc, %0;\n fma.rn.f64 %1, %4, %2, %1;\n}"
      : "+d"(acc), "+d"(vac), "+d"(one) : "d"(x), "d"(w));
}
__device__ __forceinline__ void accum_valid(float& acc, float& vac, float w, float x, float&) {
  // fp32: the NaN test and two PREDICATED accumulates, 3 instructions (the select form costs 5)
  asm("{\n .reg .pred p;\n setp.num.f32 p, %2, %2;\n @p fma.rn.f32 %0, %3, %2, %0;\n @p add.rn.f32 %1, %1, %3;\n}"
      : "+f"(acc), "+f"(vac) : "f"(x), "f"(w));
}
// acc += w * x unless x is NaN (apply_weights with skipna=False contracts fillna(0)).
__device__ __forceinline__ void accum_clean(double& acc, double w, double x) {
  asm("{\n .reg .pred p;\n .reg .b32 lo, hi;\n .reg .f64 xc;\n"
      " setp.num.f64 p, %1, %1;\n"
      " mov.b64 {lo, hi}, %1;\n selp.b32 hi, hi, 0, p;\n mov.b64 xc, {lo, hi};\n"
      " fma.rn.f64 %0, %2, xc, %0;\n}"
      : "+d"(acc) : "d"(x), "d"(w));
}
__device__ __forceinline__ void accum_clean(float& acc, float w, float x) {
  asm("{\n .reg .pred p;\n setp.num.f32 p, %1, %1;\n @p fma.rn.f32 %0, %2, %1, %0;\n}" : "+f"(acc) : "f"(x), "f"(w));
}

// Latitude contraction of the CNT taps of one output for one strip of source columns
// [col0, col0 + ncols) (mod W), 16 bytes per lane per step.  CNT is a compile-time constant: the
// CNT shared-memory loads issue back to back and the tap loop is branch-free.
template <typename T, bool SKIPNA, int CNT, int RMAX>
__device__ __forceinline__ void ring_phase1(const unsigned char* s_ring, int row_stride, int ring_slots,
                                            uint32_t slots03, uint32_t slots47, const T* __restrict__ wk,
                                            TV<T, SKIPNA>* s_tv, int W, int col0, int ncols, int pad_shift,
                                            int lane, int zero_fill) {
  constexpr int VEC = 16 / sizeof(T);
  constexpr int N = CNT > 0 ? CNT : 1;
  const unsigned char* rowp[N];
  T w[N];
  // weights of this output: one or more 16-byte loads from the staged output-major record
  {
    Pack<T, VEC> wv[(N + VEC - 1) / VEC];
#pragma unroll
    for (int i = 0; i < (CNT + VEC - 1) / VEC; ++i) wv[i] = reinterpret_cast<const Pack<T, VEC>*>(wk)[i];
#pragma unroll
    for (int i = 0; i < CNT; ++i) w[i] = wv[i / VEC].v[i % VEC];
  }
#pragma unroll
  for (int i = 0; i < CNT; ++i) {
    // ring slot of tap i: byte i of the staged per-row table + the task's rotation (see ring_row_record)
    uint32_t slot = __byte_perm(i < 4 ? slots03 : slots47, 0u, 0x4440u + static_cast<uint32_t>(i & 3));
    slot = min(slot, slot - static_cast<uint32_t>(ring_slots));  // unsigned: wraps to slot when slot < ring_slots
    rowp[i] = s_ring + slot * static_cast<uint32_t>(row_stride);
  }
  for (int q = lane * VEC; q < ncols; q += 32 * VEC) {
    int j = col0 + q;
    if (j >= W) j -= W;  // strips may wrap around the date line
    Pack<T, VEC> val[N];
#pragma unroll
    for (int i = 0; i < CNT; ++i)
      val[i] = *reinterpret_cast<const Pack<T, VEC>*>(rowp[i] + j * static_cast<int>(sizeof(T)));
    // One pass: a NaN tap is skipped by predication (value mass and valid mass), see accum_valid.
    T acc[VEC], vac[VEC], one[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) { acc[e] = T(0); vac[e] = T(0); one[e] = T(0); }
    if constexpr (SKIPNA) {
#pragma unroll
      for (int i = 0; i < CNT; ++i) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) accum_valid(acc[e], vac[e], w[i], val[i].v[e], one[e]);
      }
    } else if (zero_fill) {
#pragma unroll
      for (int i = 0; i < CNT; ++i) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) accum_clean(acc[e], w[i], val[i].v[e]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < CNT; ++i) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[e] = fma(w[i], val[i].v[e], acc[e]);
      }
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const int s = (q + e) + (pad_shift < 31 ? ((q + e) >> pad_shift) : 0);
      if constexpr (SKIPNA) {
        s_tv[s] = TV<T, true>{acc[e], vac[e]};
      } else {
        s_tv[s].t = acc[e];
      }
    }
  }
}

// The host's per-row record {need, release, kcnt, 8 x 4-bit rows-back} as the kernels keep it in shared memory:
// {need | release << 16, kcnt, 8 x 8-bit ring slot of the tap relative to the run's first row}.  The per-row path
// only adds the task's ring rotation (gbase mod ring_slots) to the packed bytes: one PRMT + one VIADDMNMX per tap
// instead of nine integer instructions.
__device__ __forceinline__ int4 ring_row_record(const int4 r, int ring_slots) {
  const int cnt = r.z > 0 ? r.z : 0;
  uint32_t lo = 0, hi = 0;
#pragma unroll
  for (int tp = 0; tp < kRingMaxTaps; ++tp) {
    const int pos = r.x - 1 - static_cast<int>((static_cast<uint32_t>(r.w) >> (4 * tp)) & 15u);
    const uint32_t sl = (tp < cnt && pos >= 0) ? static_cast<uint32_t>(pos % ring_slots) : 0u;
    if (tp < 4) lo |= sl << (8 * tp); else hi |= sl << (8 * (tp - 4));
  }
  return make_int4(r.x | (r.y << 16), r.z, static_cast<int>(lo), static_cast<int>(hi));
}

// ---------------------------------------------------------------- producer / sync roles
// producer: one thread streams the scheduled source rows of every task into the ring with TMA bulk copies
template <typename T, typename S = int32_t>
__device__ __forceinline__ void ring_producer(const T* __restrict__ x, long long xbs, long long xrs,
                                              const RingDev& ring, const S* sched_row, long long tasks,
                                              int ring_slots,
                                              uint32_t row_bytes, int row_stride, uint32_t full0, uint32_t flags0,
                                              uint32_t ring0, int n_warps) {
  uint32_t g = 0, freed = 0;  // rows issued / rows released by every consumer warp
  int slot = 0;
  const uint64_t policy = l2_policy_evict_first();
  const int W = static_cast<int>(row_bytes / sizeof(T));  // source columns of a full row
  for (long long task = blockIdx.x; task < tasks; task += gridDim.x) {
    const int panel = static_cast<int>(task % ring.n_panels);
    const long long unit = task / ring.n_panels;
    const int run = static_cast<int>(unit % ring.n_runs);
    const long long b = unit / ring.n_runs;
    const T* xb = x + b * xbs;
    const T* pb = static_cast<const T*>(ring.pole) + b * 2 * static_cast<long long>(W);
    // the panel's source columns [pc0, pc0 + pnc) modulo W: one bulk copy, or two when the range wraps
    const int pc0 = ring.panel_c0[panel], pnc = ring.panel_nc[panel];
    const int first = pnc < W - pc0 ? pnc : W - pc0;
    const uint32_t bytes1 = static_cast<uint32_t>(first) * sizeof(T);
    const uint32_t bytes2 = static_cast<uint32_t>(pnc - first) * sizeof(T);
    const int p1 = ring.run_sched[run + 1];
    for (int p = ring.run_sched[run]; p < p1; ++p) {
      const int row = sched_row[p];  // shared memory: no global load per row
      // rows >= n_src_rows are the synthetic pole rows of format_lat: full rows in the pole buffer
      const T* src = row < ring.n_src_rows
                         ? xb + static_cast<long long>(row) * xrs
                         : pb + static_cast<long long>(row - ring.n_src_rows) * W;
      uint32_t nap = 64;
      while (static_cast<int>(g - freed) >= ring_slots) {  // ring full: wait for the slowest warp
        // relaxed loads pipeline (acquire loads would serialise: 12-24 round trips per poll); one fence
        // below turns the poll into an acquire before the slot is overwritten
        uint32_t m = ld_relaxed_smem(flags0 + 4);
        for (int w = 1; w < n_warps; ++w) {
          const uint32_t d = ld_relaxed_smem(flags0 + 4 * (1 + w));
          if (static_cast<int>(d - m) < 0) m = d;
        }
        freed = m;
        if (static_cast<int>(g - freed) >= ring_slots) {
          // a full ring means the consumers are a whole ring behind: back off (the poll loop of three CTAs'
          // producers took 13 % of an SM's issue slots with a fixed 32 ns nap)
          __nanosleep(nap);
          if (nap < 512) nap *= 2;
        } else {
          fence_acq_rel_cta();
        }
      }
      const uint32_t dst = ring0 + static_cast<uint32_t>(slot) * row_stride;
      mbar_arrive_expect_tx(full0 + 8u * slot, bytes1 + bytes2);
      if (ring.l2_hint) {
        tma_load_row_hint(dst, src + pc0, bytes1, full0 + 8u * slot, policy);
        if (bytes2) tma_load_row_hint(dst + bytes1, src, bytes2, full0 + 8u * slot, policy);
      } else {
        tma_load_row(dst, src + pc0, bytes1, full0 + 8u * slot);
        if (bytes2) tma_load_row(dst + bytes1, src, bytes2, full0 + 8u * slot);
      }
      ++g;
      if (++slot == ring_slots) slot = 0;
    }
  }
}

// sync: one thread waits on the "full" mbarriers in order and publishes the landed-row count
__device__ __forceinline__ void ring_sync(const RingDev& ring, long long tasks, int ring_slots, uint32_t full0,
                                          uint32_t flags0) {
  uint32_t g = 0, parity = 0;
  int slot = 0;
  for (long long task = blockIdx.x; task < tasks; task += gridDim.x) {
    const int run = static_cast<int>((task / ring.n_panels) % ring.n_runs);
    const int n = ring.run_sched[run + 1] - ring.run_sched[run];
    for (int i = 0; i < n; ++i) {
      while (!mbar_try_wait(full0 + 8u * slot, parity)) __nanosleep(32);  // do not steal issue slots
      st_release_smem(flags0, ++g);  // rows [0, g) are in the ring and visible
      if (++slot == ring_slots) { slot = 0; parity ^= 1; }
    }
  }
}

// MAXWARPS / RMAX / MINB: launch shape.  (14 or 26, 8, 1) is the classic one-CTA-per-SM layout; (8, 4, 3) and
// (8, 8, 2) are the column-panel layouts (6 consumer warps, bands of <= 4 / <= 8 row taps, 3 / 2 CTAs per SM).
// TG: groups of 32 target columns per strip (1 classic; 2 in the panel layouts: a strip of up to 64 targets walks
// its source columns in two lane steps and its targets in two passes, halving the per-row fixed cost per target).
// (Reading the per-row tables from global memory one row ahead instead of staging them was measured and dropped:
// 3.6 -> 2.8 TB/s on 0.5 -> 1 degree.)
template <typename T, bool SKIPNA, int KW, int MAXWARPS, int RMAX, int MINB, int TG>
__global__ void __launch_bounds__(MAXWARPS * 32, MINB)
conservative_ring_kernel(const T* __restrict__ x, T* __restrict__ y, long long batch, long long xbs,
                         long long xrs, long long ybs, Band<T> cols, RingDev ring, T thr, int ring_slots,
                         int n_rows_out) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int W = cols.n_src, Wo = cols.n_out;
  const int Wp = ring.panel_cols_max;  // columns of a ring row (the whole source row when there is one panel)
  const int pad_shift = ring.pad_shift;
  const int n_warps = ring.n_warps;  // consumer warps; then one producer warp and one sync warp
  const RingLayout L = ring_layout<T>(Wp, pad_shift, SKIPNA, ring.strip_cols_max, n_warps, ring_slots);
  unsigned long long* s_full = reinterpret_cast<unsigned long long*>(smem + L.off_bar);
  uint32_t* s_flags = reinterpret_cast<uint32_t*>(smem + L.off_flags);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < ring_slots; ++s) mbar_init(&s_full[s], 1);
    fence_mbar_init();
  }
  if (tid < 32) s_flags[tid] = 0;
  // per target row tables and the row load order, staged once: no global load on the per-row paths
  // (only the RMAX weights a row of this instance can use are kept: 32 instead of 64 bytes per fp64 row)
  int4* s_krec = reinterpret_cast<int4*>(smem + L.total);
  T* s_wrec = reinterpret_cast<T*>(smem + L.total + static_cast<size_t>(n_rows_out) * 16);
  int32_t* s_sched = reinterpret_cast<int32_t*>(s_wrec + static_cast<size_t>(n_rows_out) * RMAX);
  constexpr int wstride = RMAX;  // weights per row in the staged table
  for (int i = tid; i < n_rows_out; i += blockDim.x) s_krec[i] = ring_row_record(ring.krec[i], ring_slots);
  for (int i = tid; i < n_rows_out * RMAX; i += blockDim.x)
    s_wrec[i] = static_cast<const T*>(ring.wrec)[(i / RMAX) * kRingMaxTaps + (i % RMAX)];
  for (int i = tid; i < ring.n_sched; i += blockDim.x) s_sched[i] = ring.sched_row[i];
  __syncthreads();

  const long long tasks = batch * ring.n_runs * ring.n_panels;
  const uint32_t row_bytes = static_cast<uint32_t>(W) * sizeof(T);
  const uint32_t full0 = smem_u32(s_full), flags0 = smem_u32(s_flags), ring0 = smem_u32(smem);

  // ================================================================ producer / sync warps
  if (warp == n_warps) {
    if (lane == 0)
      ring_producer<T>(x, xbs, xrs, ring, s_sched, tasks, ring_slots, row_bytes, L.row_stride, full0, flags0, ring0,
                       n_warps);
    return;
  }
  if (warp > n_warps) {
    if (warp == n_warps + 1 && lane == 0) ring_sync(ring, tasks, ring_slots, full0, flags0);
    return;
  }

  // ================================================================ consumer warps
  const T nan = quiet_nan<T>();
  using Entry = TV<T, SKIPNA>;
  Entry* s_tv = reinterpret_cast<Entry*>(smem + L.off_strip + static_cast<size_t>(warp) * L.strip_bytes);
  uint32_t gbase = 0;  // ring position of the current run's first row
  uint32_t rot = 0;    // gbase % ring_slots

  // The longitude taps of the output this lane produces stay in registers (reloaded only when the warp's strip
  // changes: a new panel, or a warp that owns several strips of the classic layout).  Indices are mapped to
  // padded slots of the strip buffer; taps beyond the output's own count get weight 0 on its first index, so
  // the longitude pass is branch-free.
  int cur = -1, l0 = 0, nl = 0, col0 = 0, ncols = 0;
  int ccnt[TG];
  T tw[TG][KW];
  int ci[TG][KW];
  auto load_strip = [&](int st, int pc0) {
    cur = st;
    l0 = ring.strip_l0[st];
    nl = ring.strip_l0[st + 1] - l0;  // host contract: <= 32 * TG
    col0 = ring.strip_c0[st];  // relative to the panel's first column
    ncols = ring.strip_nc[st];
    int abs0 = pc0 + col0;     // the strip's first source column in the full row
    if (abs0 >= W) abs0 -= W;
#pragma unroll
    for (int g = 0; g < TG; ++g) {
      const int m = lane + 32 * g;
      const int l = l0 + (m < nl ? m : 0);
      const int n = cols.cnt[l];
      ccnt[g] = (m < nl && cols.cover[l]) ? n : -1;
#pragma unroll
      for (int i = 0; i < KW; ++i) {
        const bool live = i < n && i < cols.k_max;
        int j = live ? cols.idx[i * Wo + l] : (n > 0 ? cols.idx[l] : abs0);
        j -= abs0;
        if (j < 0) j += W;
        ci[g][i] = j + (pad_shift < 31 ? (j >> pad_shift) : 0);
        tw[g][i] = live ? cols.w[i * Wo + l] : T(0);
      }
    }
  };

  for (long long task = blockIdx.x; task < tasks; task += gridDim.x) {
    const int panel = static_cast<int>(task % ring.n_panels);
    const long long unit = task / ring.n_panels;
    const int run = static_cast<int>(unit % ring.n_runs);
    const long long b = unit / ring.n_runs;
    const int k0 = ring.run_first[run], k1 = ring.run_first[run + 1];
    const int s_first = ring.panel_strip0[panel], s_end = ring.panel_strip0[panel + 1];
    const int pc0 = ring.panel_c0[panel];
    const int pw = ring.panel_nc[panel];  // strips wrap inside a full-circle panel only (pw == W)
    const uint32_t rot4 = rot * 0x01010101u;
    for (int k = k0; k < k1; ++k) {
      const int4 rec = s_krec[k];  // {need | release << 16, kcnt (-1: latitude not covered), slots 0-3, slots 4-7}
      const T* wk = s_wrec + static_cast<long long>(k) * wstride;
      const uint32_t need_g = gbase + (static_cast<uint32_t>(rec.x) & 0xffffu);
      const uint32_t rel_g = gbase + (static_cast<uint32_t>(rec.x) >> 16);
      const int cnt = rec.y > 0 ? rec.y : 0;
      const uint32_t s03 = static_cast<uint32_t>(rec.z) + rot4, s47 = static_cast<uint32_t>(rec.w) + rot4;

      // ---- wait until every row this output needs has landed in the ring
      while (static_cast<int>(ld_acquire_smem(flags0) - need_g) < 0) __nanosleep(100);

      T* yrow = y + b * ybs + static_cast<long long>(k) * Wo;
      if (s_first + warp >= s_end) {  // a panel with fewer strips than consumer warps: only report progress
        if (lane == 0) st_relaxed_smem(flags0 + 4u * (1 + warp), rel_g);
        continue;
      }
      for (int st = s_first + warp; st < s_end; st += n_warps) {
        if (st != cur) load_strip(st, pc0);
        __syncwarp();  // the previous longitude pass is done reading the strip buffer
        // ---- phase 1: latitude contraction out of the ring
#define RG_P1(C)                                                                                          \
  case C:                                                                                                 \
    ring_phase1<T, SKIPNA, C, RMAX>(smem, L.row_stride, ring_slots, s03, s47,                             \
                                    wk, s_tv, pw, col0, ncols, pad_shift, lane, ring.zero_fill);          \
    break;
        if constexpr (RMAX <= 4) {
          switch (cnt) {
            RG_P1(0) RG_P1(1) RG_P1(2) RG_P1(3)
            default:
              ring_phase1<T, SKIPNA, 4, RMAX>(smem, L.row_stride, ring_slots, s03, s47,
                                              wk, s_tv, pw, col0, ncols, pad_shift, lane, ring.zero_fill);
              break;
          }
        } else {
          switch (cnt) {
            RG_P1(0) RG_P1(1) RG_P1(2) RG_P1(3) RG_P1(4) RG_P1(5) RG_P1(6) RG_P1(7)
            default:
              ring_phase1<T, SKIPNA, 8, RMAX>(smem, L.row_stride, ring_slots, s03, s47,
                                              wk, s_tv, pw, col0, ncols, pad_shift, lane, ring.zero_fill);
              break;
          }
        }
#undef RG_P1
        __syncwarp();
        // rows that are dead after this output are no longer read once the LAST strip's latitude
        // pass is done: publish this warp's progress so the producer can recycle them early
        if (st + n_warps >= s_end && lane == 0) st_relaxed_smem(flags0 + 4u * (1 + warp), rel_g);
        // ---- phase 2: longitude contraction, normalise, mask, store
#pragma unroll
        for (int g = 0; g < TG; ++g) {
          if (lane + 32 * g < nl) {
            T out = nan;
            if (rec.y >= 0 && ccnt[g] >= 0) {
              Entry e[KW];
#pragma unroll
              for (int i = 0; i < KW; ++i) e[i] = s_tv[ci[g][i]];
              T so = T(0), sv = T(0);
#pragma unroll
              for (int i = 0; i < KW; ++i) {
                so = fma(tw[g][i], e[i].t, so);
                if constexpr (SKIPNA) sv = fma(tw[g][i], e[i].v, sv);
              }
              if constexpr (SKIPNA) out = (sv >= thr) ? normalise_div(so, sv) : nan;
              else out = so;
            }
            st_stream(yrow + l0 + lane + 32 * g, out);
          }
        }
      }
    }
    const uint32_t run_len = static_cast<uint32_t>(ring.run_sched[run + 1] - ring.run_sched[run]);
    gbase += run_len;
    rot = (rot + run_len) % static_cast<uint32_t>(ring_slots);
  }
}

// ====================================================================== lane-owned columns variant
// For the common coarsening case -- every target column takes `4 source columns + 1` contiguous taps and
// neighbouring targets are exactly 4 source columns apart (0.25 deg -> 1 deg, edge- or centre-aligned) --
// the two contractions fuse in registers: lane m of a strip's warp contracts ITS OWN 4 source columns along
// latitude (two 2-column vector loads per row tap straight out of the TMA ring), fetches the 5th column's
// {t, v} from lane m + 1 with a shuffle, and finishes the longitude contraction, normalisation, mask and
// store.  No strip buffer, one pass per target row.  Lane nl (one past the strip's last target) only produces
// the halo column.  Lanes alternate which of their two column pairs they read first (bit SWZ of the lane id)
// so that a 2-column load of 8 (fp64) / 16 (fp32) neighbouring lanes touches every bank exactly once.
// Eligibility is decided on the host (rg_row_ring_t.lane_cols == 4).
//
// Per target row and lane (round 2): 3 instructions per source element (NaN test + predicated FMA + predicated
// add, no second "redo" pass: with 1 % scattered NaNs practically every warp took it), ring slots of the row
// taps read from a per-row byte table staged at kernel start (slot = table byte + the task's ring rotation,
// one PRMT + one VIADDMNMX per tap instead of nine integer instructions), RMAX row taps at most (5 for the
// 4:1 grids: 8-tap bands use the RMAX = 8 instance; the 5-tap one has no spills).
template <typename T> struct Pair2 { T a, b; };
__device__ __forceinline__ Pair2<double> lds_pair(uint32_t addr, double) {
  Pair2<double> r;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(r.a), "=d"(r.b) : "r"(addr));
  return r;
}
__device__ __forceinline__ Pair2<float> lds_pair(uint32_t addr, float) {
  Pair2<float> r;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(r.a), "=f"(r.b) : "r"(addr));
  return r;
}

constexpr int kLaneNanPropagate = 0, kLaneNanSkip = 1, kLaneNanZero = 2;

// Latitude contraction of CNT row taps for the lane's two column pairs (A at baseA, B at baseB: ring base +
// the lane's byte offset into a ring row).  slots03 / slots47: ring slot of tap i in byte i (before wrap: the
// table byte + the task's rotation, < 2 * ring_slots).  t* = sum w x (NaN-cleaned unless MODE propagates),
// v* = valid mass (MODE skip only).
template <typename T, int MODE, int CNT, int I0 = 0, bool INIT = true>
__device__ __forceinline__ void lane_phase1(uint32_t baseA, uint32_t baseB, uint32_t row_stride,
                                            uint32_t ring_slots, uint32_t slots03, uint32_t slots47,
                                            uint32_t wk_addr, T (&t)[4], T (&v)[4], T (&one)[4]) {
  // taps [I0, CNT) of the target row; INIT = false continues the sums of an earlier call (taps [0, I0))
  constexpr int N = CNT > I0 ? CNT - I0 : 1;
  T w[N];
  if constexpr (I0 % 2 == 0) {
#pragma unroll
    for (int i = I0; i < CNT; i += 2) {  // the target row's tap weights: broadcast reads of the staged table
      const Pair2<T> p = lds_pair(wk_addr + static_cast<uint32_t>(i * sizeof(T)), T(0));
      w[i - I0] = p.a;
      if (i + 1 < CNT) w[i + 1 - I0] = p.b;
    }
  } else {
#pragma unroll
    for (int i = I0; i < CNT; ++i) {
      if constexpr (sizeof(T) == 8)
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(w[i - I0]) : "r"(wk_addr + static_cast<uint32_t>(i * sizeof(T))));
      else
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(w[i - I0]) : "r"(wk_addr + static_cast<uint32_t>(i * sizeof(T))));
    }
  }
  Pair2<T> xa[N], xb[N];
#pragma unroll
  for (int i = I0; i < CNT; ++i) {
    uint32_t slot = __byte_perm(i < 4 ? slots03 : slots47, 0u, 0x4440u + static_cast<uint32_t>(i & 3));
    slot = min(slot, slot - ring_slots);  // unsigned: wraps to slot when slot < ring_slots
    xa[i - I0] = lds_pair(baseA + slot * row_stride, T(0));
    xb[i - I0] = lds_pair(baseB + slot * row_stride, T(0));
  }
  if constexpr (INIT) {
#pragma unroll
    for (int e = 0; e < 4; ++e) { t[e] = T(0); v[e] = T(0); }
  }
#pragma unroll
  for (int i = 0; i < CNT - I0; ++i) {
    if constexpr (MODE == kLaneNanSkip) {
      accum_valid(t[0], v[0], w[i], xa[i].a, one[0]);
      accum_valid(t[1], v[1], w[i], xa[i].b, one[1]);
      accum_valid(t[2], v[2], w[i], xb[i].a, one[2]);
      accum_valid(t[3], v[3], w[i], xb[i].b, one[3]);
    } else if constexpr (MODE == kLaneNanZero) {
      accum_clean(t[0], w[i], xa[i].a);
      accum_clean(t[1], w[i], xa[i].b);
      accum_clean(t[2], w[i], xb[i].a);
      accum_clean(t[3], w[i], xb[i].b);
    } else {
      t[0] = fma(w[i], xa[i].a, t[0]);
      t[1] = fma(w[i], xa[i].b, t[1]);
      t[2] = fma(w[i], xb[i].a, t[2]);
      t[3] = fma(w[i], xb[i].b, t[3]);
    }
  }
}

// Split latitude pass (lanes kernel, one target per lane, <= 5 row taps): taps [0, 2), [2, 4) and [4, 5) are contracted
// as soon as THEIR rows have landed and released as soon as they are in registers.
// kcnt of a staged row record with the split points packed in: bits 0-7 cnt; per split point s (after tap 2: bits
// 8-15, after tap 4: bits 16-23) 4 bits "rows that may still be missing" (need - need_s) and 4 bits "rows that must
// stay" (release - release_s).  Taps in any order.  A split point that gains nothing is recorded as need offset 0.
__device__ __forceinline__ int lane_split_kcnt(const int4 raw, int points) {  // raw: the host's {need, release, kcnt, tap_back}
  const int cnt = raw.z;
  if (cnt <= 2 || cnt > 5) return cnt;
  int packed = cnt;
#pragma unroll
  for (int sp = 0; sp < 2; ++sp) {
    const int end = 2 + 2 * sp;  // taps [0, end) are done at this split point
    if (end >= cnt || sp >= points) break;
    int min_back_head = 15, max_back_tail = 0;
#pragma unroll
    for (int tp = 0; tp < 5; ++tp) {
      const int back = static_cast<int>((static_cast<uint32_t>(raw.w) >> (4 * tp)) & 15u);
      if (tp < end) min_back_head = min(min_back_head, back);
      else if (tp < cnt) max_back_tail = max(max_back_tail, back);
    }
    const int rel_s = min(raw.y, raw.x - 1 - max_back_tail);  // first row the later taps (or later targets) still need
    const int d_need = min_back_head, d_rel = raw.y - (rel_s > 0 ? rel_s : 0);
    if (d_need <= 0 || d_need > 15 || d_rel < 0 || d_rel > 15) continue;  // nothing to gain / does not fit
    packed |= (d_need | (d_rel << 4)) << (8 + 8 * sp);
  }
  return packed;
}

template <typename T, int MODE, int RMAX>
__device__ __forceinline__ void lane_phase1_dispatch(int cnt, uint32_t baseA, uint32_t baseB, uint32_t row_stride,
                                                     uint32_t ring_slots, uint32_t slots03, uint32_t slots47,
                                                     uint32_t wk, T (&t)[4], T (&v)[4], T (&one)[4]) {
#define RG_L1(C)                                                                                         \
  case C:                                                                                                \
    lane_phase1<T, MODE, C>(baseA, baseB, row_stride, ring_slots, slots03, slots47, wk, t, v, one);      \
    break;
  if constexpr (RMAX <= 5) {
    switch (cnt) {
      RG_L1(0) RG_L1(1) RG_L1(2) RG_L1(3) RG_L1(4)
      default:
        lane_phase1<T, MODE, 5>(baseA, baseB, row_stride, ring_slots, slots03, slots47, wk, t, v, one);
        break;
    }
  } else {
    switch (cnt) {
      RG_L1(0) RG_L1(1) RG_L1(2) RG_L1(3) RG_L1(4) RG_L1(5) RG_L1(6) RG_L1(7)
      default:
        lane_phase1<T, MODE, 8>(baseA, baseB, row_stride, ring_slots, slots03, slots47, wk, t, v, one);
        break;
    }
  }
#undef RG_L1
}

//
// M = 2 (rg_row_ring_t.lane_cols == 2, coarsening by 2 - 3, conservative NaN modes only): a lane still owns a
// block of 4 source columns but produces the (at most two) targets the host assigned to its block; their taps lie
// in the block or within halo_l / halo_r <= 2 columns of it, and those come from the previous / next lane with
// shuffles.  Lane 0 of a warp and the lane after its last block only produce halos.  A row wider than 12 warps
// of 30 blocks is cut into column panels (ring rows = the panel's blocks + one halo block either side); the
// launcher makes the grid a multiple of the panel count, so a CTA stays on ONE panel and the lanes' column
// offsets and weights are kernel-lifetime constants.  M = 3: the same with two halo columns either side (M = 2: one),
// M = 4: one on the left, two on the right.
template <typename T, bool SKIPNA, int MAXWARPS, int RMAX, int M = 1, int MINB = (sizeof(T) == 4 ? 2 : 1)>
__global__ void __launch_bounds__(MAXWARPS * 32, MINB)
conservative_lanes_kernel(const T* __restrict__ x, T* __restrict__ y, long long batch, long long xbs,
                          long long xrs, long long ybs, Band<T> cols, RingDev ring, T thr, int ring_slots,
                          int n_rows_out) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int W = cols.n_src, Wo = cols.n_out;
  const int n_warps = ring.n_warps;  // consumer warps; then one producer warp and one sync warp
  const RingLayout L = ring_layout<T>(ring.panel_cols_max, 31, false, 0, 0, ring_slots);  // no strip buffers
  unsigned long long* s_full = reinterpret_cast<unsigned long long*>(smem + L.off_bar);
  uint32_t* s_flags = reinterpret_cast<uint32_t*>(smem + L.off_flags);
  // per target row tables, staged once: 16-byte record + 8 tap weights (host contract: they fit)
  // (instances of <= 5 row taps keep 6 weights per row and the load order as 16-bit rows: on the headline shape that
  // is what makes room for a 19th ring row)
  constexpr int WREC = RMAX <= 5 ? kLaneWrecSmall : kRingMaxTaps;
  using Sched = typename LaneSched<RMAX>::type;
  int4* s_krec = reinterpret_cast<int4*>(smem + L.total);
  T* s_wrec = reinterpret_cast<T*>(smem + L.total + static_cast<size_t>(n_rows_out) * 16);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < ring_slots; ++s) mbar_init(&s_full[s], 1);
    fence_mbar_init();
  }
  if (tid < 32) s_flags[tid] = 0;
  // SPLIT (one target per lane, <= 5 row taps): the latitude pass of a target row runs in up to three parts -- taps
  // [0, 2), [2, 4), [4, 5), each as soon as ITS rows have landed, released as soon as they are in registers -- so that
  // ring slots are not held by rows waiting for the last row of their target row: more bytes in flight for the same
  // ring (the kernel's rate follows the SM clock, i.e. it is bound by bytes in flight over latency)
  constexpr bool SPLIT = M == 1 && RMAX <= 5;
  for (int i = tid; i < n_rows_out; i += blockDim.x) {
    const int4 raw = ring.krec[i];
    int4 rec = ring_row_record(raw, ring_slots);
    if constexpr (SPLIT) {
      if (ring.lane_split) rec.y = lane_split_kcnt(raw, ring.lane_split);
    }
    s_krec[i] = rec;
  }
  for (int i = tid; i < n_rows_out * WREC; i += blockDim.x)
    s_wrec[i] = static_cast<const T*>(ring.wrec)[(i / WREC) * kRingMaxTaps + i % WREC];
  Sched* s_sched = reinterpret_cast<Sched*>(s_wrec + static_cast<size_t>(n_rows_out) * WREC);
  for (int i = tid; i < ring.n_sched; i += blockDim.x) s_sched[i] = static_cast<Sched>(ring.sched_row[i]);
  __syncthreads();

  const long long tasks = batch * ring.n_runs * ring.n_panels;  // (M = 1: one panel)
  const uint32_t row_bytes = static_cast<uint32_t>(W) * sizeof(T);
  const uint32_t full0 = smem_u32(s_full), flags0 = smem_u32(s_flags), ring0 = smem_u32(smem);

  if (warp == n_warps) {
    if (lane == 0)
      ring_producer<T, Sched>(x, xbs, xrs, ring, s_sched, tasks, ring_slots, row_bytes, L.row_stride, full0, flags0,
                              ring0, n_warps);
    return;
  }
  if (warp > n_warps) {
    if (warp == n_warps + 1 && lane == 0) ring_sync(ring, tasks, ring_slots, full0, flags0);
    return;
  }

  // ================================================================ consumer warps (strip = warp)
  const T nan = quiet_nan<T>();
  // M = 2: gridDim.x is a multiple of n_panels, so task % n_panels is the same for every task of this CTA
  const int panel = M >= 2 ? static_cast<int>(blockIdx.x % static_cast<unsigned>(ring.n_panels)) : 0;
  const int strip = (M >= 2 ? ring.panel_strip0[panel] : 0) + warp;
  // host contract: strips of a panel <= n_warps
  const bool has_strip = M >= 2 ? strip < ring.panel_strip0[panel + 1] : warp < ring.n_strips;
  const int l0 = has_strip ? ring.strip_l0[strip] : 0;
  const int nl = has_strip ? ring.strip_l0[strip + 1] - l0 : 0;  // host contract: <= 31 (M = 2: <= 60)
  const int m = lane < nl ? lane : (nl > 0 ? nl - 1 : 0);
  const int l = l0 + m;
  // first source column of this lane's group of 4 (the halo lane continues one group further)
  int cfirst = (has_strip && nl > 0) ? cols.idx[l] + 4 * ((lane < nl ? lane : nl) - m) : 0;
  int b0 = 0, nb = 0, cabs = 0;  // M = 2: first block / blocks of the strip (<= 30), this lane's first own column
  bool wrap_cols = true;
  if constexpr (M >= 2) {
    b0 = has_strip ? ring.strip_b0[strip] : 0;
    nb = has_strip ? ring.strip_b0[strip + 1] - b0 : 0;
    const int blk = (lane <= nb ? lane : nb + 1) - 1;  // -1: left halo block, nb: right halo block
    cabs = has_strip ? ring.strip_c0[strip] + 4 * blk : 0;
    // column inside the ring row: rows hold the panel's columns [pc0, pc0 + pnc) (mod W)
    const int pc0 = ring.panel_c0[panel];
    wrap_cols = ring.panel_nc[panel] == W;  // a full-circle row: blocks may straddle its end
    cfirst = has_strip ? (((cabs - pc0) % W) + W) % W : 0;
  }
  constexpr int SWZ = sizeof(T) == 8 ? 2 : 3;
  const int swz = (lane >> SWZ) & 1;
  int jA = cfirst + 2 * swz, jB = cfirst + 2 * (1 - swz);
  if (wrap_cols) {
    jA %= W;
    jB %= W;
  }
  const uint32_t baseA = ring0 + static_cast<uint32_t>(jA) * sizeof(T);
  const uint32_t baseB = ring0 + static_cast<uint32_t>(jB) * sizeof(T);
  // longitude taps of this lane's target: weights in register order {A.a, A.b, B.a, B.b, neighbour}
  const int ccnt = (has_strip && lane < nl && cols.cover[l]) ? cols.cnt[l] : -1;
  T tw[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) tw[i] = (i < ccnt && i < cols.k_max) ? cols.w[i * Wo + l] : T(0);
  const T twA0 = swz ? tw[2] : tw[0], twA1 = swz ? tw[3] : tw[1];
  const T twB0 = swz ? tw[0] : tw[2], twB1 = swz ? tw[1] : tw[3];
  const T tw4 = tw[4];
  // M = 2: weights of the lane's targets over the window {2 left halo columns, own 4 in register order, 2 right}
  T tw2[M >= 2 ? 2 : 1][8];
  bool have2[2] = {false, false};
  int ntg = 0, tgt0 = 0;
  bool pair_store = false;
  if constexpr (M >= 2) {
    if (has_strip && lane >= 1 && lane <= nb) {
      tgt0 = ring.block_l0[b0 + lane - 1];
      ntg = ring.block_l0[b0 + lane] - tgt0;  // host contract: 0..2
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const bool live = j < ntg;
      const int lj = live ? tgt0 + j : 0;
      have2[j] = live && cols.cover[lj] != 0;
      T nat[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) nat[q] = T(0);
      if (have2[j]) {
        const int cj = cols.cnt[lj];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          if (i < cj && i < cols.k_max) {
            const int q = (((cols.idx[i * Wo + lj] - (cabs - 2)) % W) + W) % W;
            const T wv = cols.w[i * Wo + lj];
#pragma unroll
            for (int qq = 0; qq < 8; ++qq)
              if (qq == q) nat[qq] = wv;
          }
        }
      }
      tw2[j][0] = nat[0];
      tw2[j][1] = nat[1];
      tw2[j][2] = swz ? nat[4] : nat[2];
      tw2[j][3] = swz ? nat[5] : nat[3];
      tw2[j][4] = swz ? nat[2] : nat[4];
      tw2[j][5] = swz ? nat[3] : nat[5];
      tw2[j][6] = nat[6];
      tw2[j][7] = nat[7];
    }
    // both targets of a lane in one 2-element store where that address is aligned
    pair_store = ntg == 2 && (Wo % 2 == 0) && (ybs % 2 == 0) && (tgt0 % 2 == 0) &&
                 reinterpret_cast<uintptr_t>(y) % (2 * sizeof(T)) == 0;
  }
  constexpr bool halo_l2 = M == 3, halo_r2 = M == 3 || M == 4;  // (compile time: a predicated fp64 FMA costs three instructions)
  // zero-weight taps must not see the data (0 * NaN) -- only where NaNs survive the latitude pass
  const bool padded = !SKIPNA && !ring.zero_fill && ccnt >= 0 && ccnt < 5;
  const bool zero_fill = !SKIPNA && ring.zero_fill;
  const bool writer = M >= 2 ? ntg > 0 : lane < nl;
  T* const ylane = y + (M >= 2 ? tgt0 : l0 + lane);

  const uint32_t wrec0 = smem_u32(s_wrec);
  const uint32_t row_stride = static_cast<uint32_t>(L.row_stride), slots = static_cast<uint32_t>(ring_slots);
  // 0 / 1 factors of accum_valid, one register pair per column for the whole kernel.  Their low words are zeros
  // READ from shared memory (an unused flag word): a zero the compiler cannot see through stays in its register
  // instead of being rematerialised in front of every use.
  T one[4];
  {
    uint32_t z;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(z) : "r"(flags0 + 124u));
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if constexpr (sizeof(T) == 8) one[e] = __hiloint2double(static_cast<int>(z), static_cast<int>(z));
      else one[e] = __uint_as_float(z);
    }
  }
  uint32_t gbase = 0;  // ring position of the current run's first row
  uint32_t rot = 0;    // gbase % ring_slots

  for (long long task = blockIdx.x; task < tasks; task += gridDim.x) {
    const long long unit = M >= 2 ? task / ring.n_panels : task;
    const int run = static_cast<int>(unit % ring.n_runs);
    const long long b = unit / ring.n_runs;
    const int k0 = ring.run_first[run], k1 = ring.run_first[run + 1];
    const uint32_t rot4 = rot * 0x01010101u;
    T* yrow = ylane + b * ybs + static_cast<long long>(k0) * Wo;
    for (int k = k0; k < k1; ++k, yrow += Wo) {
      const int4 rec = s_krec[k];  // {need | release << 16, kcnt (-1: latitude not covered), slots 0-3, slots 4-7}
      const uint32_t need_g = gbase + (static_cast<uint32_t>(rec.x) & 0xffffu);
      const int cnt = rec.y > 0 ? (rec.y & 0xff) : 0;
      const uint32_t wk = wrec0 + static_cast<uint32_t>(k) * (WREC * sizeof(T));
      const uint32_t s03 = static_cast<uint32_t>(rec.z) + rot4, s47 = static_cast<uint32_t>(rec.w) + rot4;
      T t[4], v[4];
      const bool split = SPLIT && rec.y > 0xff;  // (warp-uniform)
      if (split) {
        // MODE is a compile-time constant per instance except for the zero-fill flag of the !SKIPNA kernels
#define RG_PART(CNTV, I0V, INITV)                                                                                   \
  do {                                                                                                                \
    if constexpr (SKIPNA) {                                                                                           \
      lane_phase1<T, kLaneNanSkip, CNTV, I0V, INITV>(baseA, baseB, row_stride, slots, s03, s47, wk, t, v, one);        \
    } else {                                                                                                          \
      if (zero_fill) lane_phase1<T, kLaneNanZero, CNTV, I0V, INITV>(baseA, baseB, row_stride, slots, s03, s47, wk, t, v, one); \
      else lane_phase1<T, kLaneNanPropagate, CNTV, I0V, INITV>(baseA, baseB, row_stride, slots, s03, s47, wk, t, v, one);      \
    }                                                                                                                 \
  } while (0)
        const uint32_t rel_g = gbase + (static_cast<uint32_t>(rec.x) >> 16);
        const uint32_t sp1 = (static_cast<uint32_t>(rec.y) >> 8) & 0xffu, sp2 = (static_cast<uint32_t>(rec.y) >> 16) & 0xffu;
        if (sp1 & 15u) {  // taps [0, 2) ahead of the rest
          while (static_cast<int>(ld_acquire_smem(flags0) - (need_g - (sp1 & 15u))) < 0) __nanosleep(100);
          RG_PART(2, 0, true);
          __syncwarp();
          if (lane == 0) st_relaxed_smem(flags0 + 4u * (1 + warp), rel_g - (sp1 >> 4));
          if (sp2 & 15u) {  // taps [2, 4) ahead of tap 4 (cnt == 5)
            while (static_cast<int>(ld_acquire_smem(flags0) - (need_g - (sp2 & 15u))) < 0) __nanosleep(100);
            RG_PART(4, 2, false);
            __syncwarp();
            if (lane == 0) st_relaxed_smem(flags0 + 4u * (1 + warp), rel_g - (sp2 >> 4));
            while (static_cast<int>(ld_acquire_smem(flags0) - need_g) < 0) __nanosleep(100);
            RG_PART(5, 4, false);
          } else {
            while (static_cast<int>(ld_acquire_smem(flags0) - need_g) < 0) __nanosleep(100);
            if (cnt == 3) RG_PART(3, 2, false);
            else if (cnt == 4) RG_PART(4, 2, false);
            else RG_PART(5, 2, false);
          }
        } else {  // only the second split point: taps [0, 4), then tap 4
          while (static_cast<int>(ld_acquire_smem(flags0) - (need_g - (sp2 & 15u))) < 0) __nanosleep(100);
          RG_PART(4, 0, true);
          __syncwarp();
          if (lane == 0) st_relaxed_smem(flags0 + 4u * (1 + warp), rel_g - (sp2 >> 4));
          while (static_cast<int>(ld_acquire_smem(flags0) - need_g) < 0) __nanosleep(100);
          RG_PART(5, 4, false);
        }
#undef RG_PART
      } else {
        while (static_cast<int>(ld_acquire_smem(flags0) - need_g) < 0) __nanosleep(100);
        if constexpr (SKIPNA) {
          lane_phase1_dispatch<T, kLaneNanSkip, RMAX>(cnt, baseA, baseB, row_stride, slots, s03, s47, wk, t, v, one);
        } else {
          if (zero_fill)
            lane_phase1_dispatch<T, kLaneNanZero, RMAX>(cnt, baseA, baseB, row_stride, slots, s03, s47, wk, t, v, one);
          else
            lane_phase1_dispatch<T, kLaneNanPropagate, RMAX>(cnt, baseA, baseB, row_stride, slots, s03, s47, wk, t, v, one);
        }
      }
      // every value this warp needs from the rows that die after this output is now in registers
      __syncwarp();
      if (lane == 0) st_relaxed_smem(flags0 + 4u * (1 + warp), gbase + (static_cast<uint32_t>(rec.x) >> 16));

      if constexpr (M >= 2) {
        // halos: the previous lane's last (two) columns, the next lane's first (two)
        const T t_n0 = swz ? t[2] : t[0], t_n1 = swz ? t[3] : t[1], t_n2 = swz ? t[0] : t[2], t_n3 = swz ? t[1] : t[3];
        const T tl1 = __shfl_up_sync(0xffffffffu, t_n3, 1), tr0 = __shfl_down_sync(0xffffffffu, t_n0, 1);
        T tl0 = T(0), tr1 = T(0), vl0 = T(0), vl1 = T(0), vr0 = T(0), vr1 = T(0);
        if constexpr (halo_l2) tl0 = __shfl_up_sync(0xffffffffu, t_n2, 1);
        if constexpr (halo_r2) tr1 = __shfl_down_sync(0xffffffffu, t_n1, 1);
        if constexpr (SKIPNA) {
          const T v_n0 = swz ? v[2] : v[0], v_n1 = swz ? v[3] : v[1], v_n2 = swz ? v[0] : v[2], v_n3 = swz ? v[1] : v[3];
          vl1 = __shfl_up_sync(0xffffffffu, v_n3, 1);
          vr0 = __shfl_down_sync(0xffffffffu, v_n0, 1);
          if constexpr (halo_l2) vl0 = __shfl_up_sync(0xffffffffu, v_n2, 1);
          if constexpr (halo_r2) vr1 = __shfl_down_sync(0xffffffffu, v_n1, 1);
        }
        if (writer) {
          T out[2];
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            out[j] = nan;
            if (rec.y >= 0 && have2[j]) {
              // two short chains per sum (a dependent fp64 FMA waits ~10 cycles)
              T so = tw2[j][1] * tl1, so2 = tw2[j][4] * t[2];
              so = fma(tw2[j][2], t[0], so);
              so2 = fma(tw2[j][5], t[3], so2);
              so = fma(tw2[j][3], t[1], so);
              so2 = fma(tw2[j][6], tr0, so2);
              if constexpr (halo_l2) so = fma(tw2[j][0], tl0, so);
              if constexpr (halo_r2) so2 = fma(tw2[j][7], tr1, so2);
              so += so2;
              if constexpr (SKIPNA) {
                T sv = tw2[j][1] * vl1, sv2 = tw2[j][4] * v[2];
                sv = fma(tw2[j][2], v[0], sv);
                sv2 = fma(tw2[j][5], v[3], sv2);
                sv = fma(tw2[j][3], v[1], sv);
                sv2 = fma(tw2[j][6], vr0, sv2);
                if constexpr (halo_l2) sv = fma(tw2[j][0], vl0, sv);
                if constexpr (halo_r2) sv2 = fma(tw2[j][7], vr1, sv2);
                sv += sv2;
                out[j] = (sv >= thr) ? normalise_div(so, sv) : nan;
              } else {
                out[j] = so;
              }
            }
          }
          if (pair_store) {
            st_stream2(yrow, out[0], out[1]);
          } else {
            st_stream(yrow, out[0]);
            if (ntg == 2) st_stream(yrow + 1, out[1]);
          }
        }
        continue;
      }
      // the 5th tap is the first column of the next lane's group
      const T t_first = swz ? t[2] : t[0];
      const T tn = __shfl_down_sync(0xffffffffu, t_first, 1);
      T vn = T(0);
      if constexpr (SKIPNA) {
        const T v_first = swz ? v[2] : v[0];
        vn = __shfl_down_sync(0xffffffffu, v_first, 1);
      }
      if (writer) {
        T out = nan;
        if (rec.y >= 0 && ccnt >= 0) {
          T e0 = t[0], e1 = t[1], e2 = t[2], e3 = t[3], e4 = tn;
          if (padded) {
            e0 = twA0 != T(0) ? e0 : T(0);
            e1 = twA1 != T(0) ? e1 : T(0);
            e2 = twB0 != T(0) ? e2 : T(0);
            e3 = twB1 != T(0) ? e3 : T(0);
            e4 = tw4 != T(0) ? e4 : T(0);
          }
          T so = twA0 * e0;
          so = fma(twA1, e1, so);
          so = fma(twB0, e2, so);
          so = fma(twB1, e3, so);
          so = fma(tw4, e4, so);
          if constexpr (SKIPNA) {
            T sv = twA0 * v[0];
            sv = fma(twA1, v[1], sv);
            sv = fma(twB0, v[2], sv);
            sv = fma(twB1, v[3], sv);
            sv = fma(tw4, vn, sv);
            out = (sv >= thr) ? normalise_div(so, sv) : nan;
          } else {
            out = so;
          }
        }
        st_stream(yrow, out);
      }
    }
    const uint32_t run_len = static_cast<uint32_t>(ring.run_sched[run + 1] - ring.run_sched[run]);
    gbase += run_len;
    rot = (rot + run_len) % slots;
  }
}

}  // namespace rg
