// Fused conservative regridding for sm_100a.
//
// One pass over the source stack: for a (slab b, target row k) task a CTA
//   phase 1  streams the <= k_max source rows of the latitude band with 16-byte
//            read-only loads, contracts them per source column into shared memory
//            (value mass t[j] and, for skipna, valid mass v[j]);
//   phase 2  contracts t / v along longitude out of shared memory with the column
//            band (whose indices already contain the reference's +-360 shift and
//            wrap padding), divides, applies the nan_threshold and coverage masks and
//            writes the Wo outputs coalesced.
// No intermediate touches HBM.  Replaces apply_weights + masks of the reference
// (src/xarray_regrid/methods/conservative.py:181-208, 161-164).
//
// Roofline: HBM.  Algorithmic bytes per task = (band rows * W + Wo) * sizeof(T); flops are
// ~0.4 per byte, far below the fp64 ridge, so tensor cores are not used.

#include "common.cuh"
#include <cstdlib>

#include "conservative_ring.cuh"
#include "pipeline.cuh"
#include "upsample.cuh"

namespace rg {

constexpr int kConsThreads = 256;

// What a NaN in the source does (conservative_launch's nan_mode):
//   kNanPropagate  poisons every output it touches, weight 0 included (interpolation: scipy's w*y sums)
//   kNanSkip       apply_weights(skipna=True): fillna(0) + valid-mass normalisation + threshold
//   kNanZeroFill   apply_weights(skipna=False): the reference contracts da.fillna(0) unconditionally
//                  (methods/conservative.py:196-198), so a NaN counts as 0 and the result is finite
constexpr int kNanPropagate = 0, kNanSkip = 1, kNanZeroFill = 2;

// RC = independent row loads in flight per thread (the row taps are walked in chunks of RC; bands with <= 4 row
// taps use RC = 4: the unrolled, predicated chunk body is half as long).
template <typename T, bool SKIPNA, int VEC, int RC>
__global__ void __launch_bounds__(kConsThreads)
conservative_kernel(const T* __restrict__ x, T* __restrict__ y, long long batch,
                    long long xbs, long long xrs, long long ybs, Band<T> rows, Band<T> cols,
                    T thr, const T* __restrict__ pole, int w_pad, int zero_fill) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* s_t = reinterpret_cast<T*>(smem_raw);
  T* s_v = s_t + w_pad;  // valid only when SKIPNA
  long long* s_off = reinterpret_cast<long long*>(s_t + (SKIPNA ? 2 : 1) * w_pad);
  T* s_w = reinterpret_cast<T*>(s_off + rows.k_max);

  const int H = rows.n_src, W = cols.n_src, Ho = rows.n_out, Wo = cols.n_out;
  const long long tasks = batch * Ho;
  const T nan = quiet_nan<T>();

  for (long long task = blockIdx.x; task < tasks; task += gridDim.x) {
    const int k = static_cast<int>(task % Ho);
    const long long b = task / Ho;
    T* yrow = y + b * ybs + static_cast<long long>(k) * Wo;

    if (!rows.cover[k]) {  // CTA-uniform: target latitude outside the source range
      for (int l = threadIdx.x; l < Wo; l += kConsThreads) st_stream(yrow + l, nan);
      continue;
    }

    const int cnt = rows.cnt[k];
    if (threadIdx.x < cnt) {
      const int r = rows.idx[threadIdx.x * Ho + k];
      // r >= H addresses the synthetic pole rows (value pole[b, r - H] for every column)
      s_off[threadIdx.x] = r < H ? static_cast<long long>(r) * xrs : -static_cast<long long>(r - H) - 1;
      s_w[threadIdx.x] = rows.w[threadIdx.x * Ho + k];
    }
    __syncthreads();

    // ---- phase 1: latitude contraction, one source column group per thread ----------
    const T* xb = x + b * xbs;
    for (int j = threadIdx.x * VEC; j < W; j += kConsThreads * VEC) {
      T acc[VEC], vac[VEC];
#pragma unroll
      for (int e = 0; e < VEC; ++e) { acc[e] = T(0); vac[e] = T(0); }
      const bool full = (j + VEC <= W);
      for (int r0 = 0; r0 < cnt; r0 += RC) {
        Pack<T, VEC> val[RC];
#pragma unroll
        for (int i = 0; i < RC; ++i) {
          if (r0 + i < cnt) {
            const long long off = s_off[r0 + i];
            if (off >= 0) {
              if (full) {
                val[i] = ld_stream<T, VEC>(xb + off + j);
              } else {
#pragma unroll
                for (int e = 0; e < VEC; ++e)
                  val[i].v[e] = (j + e < W) ? ld_stream<T, 1>(xb + off + j + e).v[0] : T(0);
              }
            } else {
              const T pv = pole[(b * 2 + (-off - 1)) * static_cast<long long>(W)];
#pragma unroll
              for (int e = 0; e < VEC; ++e) val[i].v[e] = pv;
            }
          }
        }
#pragma unroll
        for (int i = 0; i < RC; ++i) {
          if (r0 + i < cnt) {
            const T w = s_w[r0 + i];
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
              const T xv = val[i].v[e];
              if (SKIPNA) {
                const bool ok = (xv == xv);
                acc[e] = fma(w, ok ? xv : T(0), acc[e]);
                vac[e] += ok ? w : T(0);
              } else {
                // zero_fill: apply_weights(skipna=False) contracts da.fillna(0) (conservative.py:196-198);
                // otherwise (interpolation callers) a NaN tap poisons the output
                acc[e] = fma(w, (zero_fill && xv != xv) ? T(0) : xv, acc[e]);
              }
            }
          }
        }
      }
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        if (j + e < W) {
          s_t[j + e] = acc[e];
          if (SKIPNA) s_v[j + e] = vac[e];
        }
      }
    }
    __syncthreads();

    // ---- phase 2: longitude contraction out of shared memory, normalise, mask, store -
    for (int l = threadIdx.x; l < Wo; l += kConsThreads) {
      T out = nan;
      if (cols.cover[l]) {
        const int c = cols.cnt[l];
        T so = T(0), sv = T(0);
        for (int t = 0; t < c; ++t) {
          const int j = cols.idx[t * Wo + l];
          const T w = cols.w[t * Wo + l];
          so = fma(w, s_t[j], so);
          if (SKIPNA) sv = fma(w, s_v[j], sv);
        }
        if (SKIPNA) {
          out = (sv >= thr) ? so / sv : nan;
        } else {
          out = so;
        }
      }
      st_stream(yrow + l, out);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ pole rows (format_lat)
template <typename T>
__global__ void __launch_bounds__(256)
row_nanmean_kernel(const T* __restrict__ x, T* __restrict__ pole, long long batch, int W,
                   long long xbs, long long xrs, int south_row, int north_row) {
  __shared__ double s_sum[8];
  __shared__ long long s_cnt[8];
  const long long b = blockIdx.x >> 1;
  const int which = blockIdx.x & 1;
  const T* row = x + b * xbs + static_cast<long long>(which ? north_row : south_row) * xrs;
  double sum = 0.0;
  long long n = 0;
  for (int j = threadIdx.x; j < W; j += 256) {
    const T v = row[j];
    if (v == v) { sum += static_cast<double>(v); ++n; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    n += __shfl_xor_sync(0xffffffffu, n, o);
  }
  if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5] = sum; s_cnt[threadIdx.x >> 5] = n; }
  __syncthreads();
  // every thread folds the 8 partials in the same order, then the mean is written as a FULL row of W
  // copies: the pole rows are ordinary rows for the TMA row-ring kernels (and element 0 for the generic one)
  double ts = 0.0;
  long long tn = 0;
  for (int i = 0; i < 8; ++i) { ts += s_sum[i]; tn += s_cnt[i]; }
  const T mean = tn > 0 ? static_cast<T>(ts / static_cast<double>(tn)) : quiet_nan<T>();
  T* out = pole + (b * 2 + which) * static_cast<long long>(W);
  for (int j = threadIdx.x; j < W; j += 256) out[j] = mean;
}

template <typename T>
int row_nanmean_launch(const T* x, T* pole, int64_t batch, int32_t H, int32_t W, int64_t xbs,
                       int64_t xrs, int32_t south_row, int32_t north_row, cudaStream_t stream) {
  if (!x || !pole) return RG_ERR_NULL;
  if (batch < 0 || H <= 0 || W <= 0 || xrs < W) return RG_ERR_SHAPE;
  if (south_row < 0 || south_row >= H || north_row < 0 || north_row >= H) return RG_ERR_SHAPE;
  if (batch == 0) return RG_OK;
  if (batch * 2 > 0x7fffffffLL) return RG_ERR_SHAPE;
  row_nanmean_kernel<T><<<static_cast<unsigned>(batch * 2), 256, 0, stream>>>(x, pole, batch, W, xbs, xrs,
                                                                              south_row, north_row);
  return static_cast<int>(cudaGetLastError());
}

// --------------------------------------------------------------------------- launcher
template <typename T>
int conservative_launch(const T* x, T* y, int64_t batch, int64_t xbs, int64_t xrs, int64_t ybs,
                        const rg_band_t* rows_c, const rg_band_t* cols_c, int32_t nan_mode, T thr,
                        const T* pole, const rg_row_ring_t* ring_c, cudaStream_t stream) {
  if (!x || !y) return RG_ERR_NULL;
  if (nan_mode < kNanPropagate || nan_mode > kNanZeroFill) return RG_ERR_UNSUPPORTED;
  const int32_t skipna = nan_mode == kNanSkip;
  const int zero_fill = nan_mode == kNanZeroFill;
  Band<T> rows, cols;
  int rc = make_band<T>(rows_c, &rows);
  if (rc != RG_OK) return rc;
  rc = make_band<T>(cols_c, &cols);
  if (rc != RG_OK) return rc;
  if (batch < 0) return RG_ERR_SHAPE;
  if (batch == 0) return RG_OK;
  const int W = cols.n_src;
  if (xrs < W || xbs < 0 || ybs < static_cast<int64_t>(rows.n_out) * cols.n_out) return RG_ERR_SHAPE;
  if (rows.k_max > kConsThreads) return RG_ERR_UNSUPPORTED;

  constexpr int kMaxVec = 16 / sizeof(T);
  const bool aligned = (reinterpret_cast<uintptr_t>(x) % 16 == 0) && (xbs % kMaxVec == 0) &&
                       (xrs % kMaxVec == 0);
  const int w_pad = (W + kMaxVec - 1) / kMaxVec * kMaxVec;
  const size_t smem = static_cast<size_t>(skipna ? 2 : 1) * w_pad * sizeof(T) +
                      static_cast<size_t>(rows.k_max) * (sizeof(long long) + sizeof(T));

  DeviceInfo dev;
  rc = device_info(&dev);
  if (rc != RG_OK) return rc;

  // ---- interpolation onto a much finer grid (write-bound): shared-memory upsampling kernel
  if (nan_mode == kNanPropagate && !pole) {
    rc = upsample_try_launch<T>(x, y, batch, xbs, xrs, ybs, rows, cols, dev, stream);
    if (rc != RG_ERR_UNSUPPORTED) return rc;
  }

  // ---- fast path: TMA row-ring kernels (need 16-byte aligned rows, <= 8 taps per
  //      axis, a ring that fits shared memory)
  // lane_cols == 2: the register-resident kernel in its block layout (up to two targets per lane, <= 30 blocks
  // per warp, conservative NaN modes only).  That layout has no two-phase form: anything else goes to the generic
  // kernel.
  const bool lanes2 = ring_c && ring_c->lane_cols == 2 && cols.k_max <= 6 && rows.k_max <= 5 && rows.n_src + 2 < 65536 &&
                      (skipna || zero_fill) && ring_c->strip_targets_max <= 60 && ring_c->strip_b0 &&
                      ring_c->block_l0 && ring_c->halo_l >= 1 && ring_c->halo_l <= 2 && ring_c->halo_r >= 1 &&
                      ring_c->halo_r <= 2 && ring_c->n_panels >= 1 && ring_c->panel_strip0 && ring_c->panel_c0 &&
                      ring_c->panel_nc && ring_c->n_warps + 2 <= 14 && W % 2 == 0 &&
                      ring_c->panel_cols_max <= W && (ring_c->panel_cols_max * sizeof(T)) % 16 == 0 &&
                      (ring_c->n_panels == 1 || W % 4 == 0);
  if (ring_c && ring_c->lane_cols == 2 && !lanes2) ring_c = nullptr;
  if (ring_c && aligned && (static_cast<size_t>(W) * sizeof(T)) % 16 == 0 &&
      (!pole || reinterpret_cast<uintptr_t>(pole) % 16 == 0) &&
      rows.k_max <= kRingMaxTaps && cols.k_max <= 8 && ring_c->n_runs > 0 && ring_c->n_strips > 0 &&
      ring_c->n_sched < 65536 &&  // the staged per-row record packs need / release into 16 bits each
      ring_c->n_warps >= 1 && ring_c->n_warps <= kRingMaxWarps && ring_c->max_live <= kRingMaxLive &&
      (ring_c->strip_targets_max <= 32 || lanes2 ||
       (ring_c->strip_targets_max <= 64 && ring_c->n_warps <= 6 && cols.k_max >= 3)) &&
      ring_c->sched_row && ring_c->krec && ring_c->wrec && ring_c->run_first && ring_c->run_sched &&
      ring_c->strip_l0 && ring_c->strip_c0 && ring_c->strip_nc) {
    const int kw = cols.k_max <= 1 ? 1 : cols.k_max <= 2 ? 2 : cols.k_max <= 3 ? 3 : cols.k_max <= 5 ? 5 : 8;
    // column panels (a missing table = one panel covering the whole row)
    const bool have_panels = ring_c->n_panels >= 1 && ring_c->panel_strip0 && ring_c->panel_c0 && ring_c->panel_nc;
    const int n_panels = have_panels ? ring_c->n_panels : 1;
    const int wp = have_panels ? ring_c->panel_cols_max : W;  // columns of a ring row
    // lane-owned columns variant (4 source columns + 1 per target, see conservative_ring.cuh): no strip buffers
    // (the staged per-row record packs need / release into 16 bits each)
    const bool lanes = lanes2 || (ring_c->lane_cols == 4 && cols.k_max <= 5 && ring_c->n_strips <= ring_c->n_warps &&
                                  ring_c->n_warps + 2 <= 14 && W % 2 == 0 && n_panels == 1 && wp == W);
    // panel layouts of the two-phase kernel: 6 consumer warps, several CTAs per SM
    const bool panel_cta = !lanes && have_panels && ring_c->n_warps <= 6 && kw >= 3;
    const bool panel_tg2 = panel_cta && ring_c->strip_targets_max > 32;
    // (register-resident kernel: 14 warps of fp64 fill an SM's registers; rows of fewer consumer warps leave
    // room for more CTAs)
    using LanesKernel = void (*)(const T*, T*, long long, long long, long long, long long, Band<T>, RingDev, T, int, int);
    LanesKernel lk = nullptr;
    int lane_ctas = 1;
    if (lanes) {
      // block layout: M = 2 / 3 / 4 = halo columns (1, 1) / (2, 2) / (1, 2); narrow rows (<= 3 consumer warps per
      // CTA) run four (fp64) / five (fp32) CTAs per SM -- the register cap of (160 threads, 4 CTAs) is 96
      const bool narrow = ring_c->n_warps <= 3;
      constexpr int kNarrowCtas = 4;  // register cap 96; fp32 (72 registers) then fits five: 3.65 TB/s on 1 -> 2.5 deg,
                                      // an instance capped at 64 registers for six CTAs 3.44 TB/s
      const int m = !lanes2 ? 1 : ring_c->halo_l > 1 ? 3 : ring_c->halo_r > 1 ? 4 : 2;
#define RG_BLOCKS(MV)                                                                                          \
  (narrow ? (skipna ? conservative_lanes_kernel<T, true, 5, 5, MV, kNarrowCtas> : conservative_lanes_kernel<T, false, 5, 5, MV, kNarrowCtas>) \
          : (skipna ? conservative_lanes_kernel<T, true, 14, 5, MV> : conservative_lanes_kernel<T, false, 14, 5, MV>))
      lk = m == 3 ? RG_BLOCKS(3) : m == 4 ? RG_BLOCKS(4) : m == 2 ? RG_BLOCKS(2)
#undef RG_BLOCKS
           : (rows.k_max <= 5 && rows.n_src + 2 < 65536)
               ? (skipna ? conservative_lanes_kernel<T, true, 14, 5> : conservative_lanes_kernel<T, false, 14, 5>)
               : (skipna ? conservative_lanes_kernel<T, true, 14, 8> : conservative_lanes_kernel<T, false, 14, 8>);
      cudaFuncAttributes fa;
      RG_CUDA(cudaFuncGetAttributes(&fa, lk));
      const int regs = (fa.numRegs + 7) / 8 * 8;  // allocated per warp in units of 8 registers per thread
      lane_ctas = 65536 / ((ring_c->n_warps + 2) * 32 * regs);
    }
    const int min_ctas = lanes ? (lane_ctas > 8 ? 8 : lane_ctas < 1 ? 1 : lane_ctas)
                         : panel_cta ? ((rows.k_max <= 4 && !(panel_tg2 && kw >= 5)) ? 3 : 2) : 1;
    const RingLayout probe =
        lanes ? ring_layout<T>(wp, 31, false, 0, 0, 0)
              : ring_layout<T>(wp, ring_c->pad_shift, skipna != 0, ring_c->strip_cols_max, ring_c->n_warps, 0);
    // all variants stage the per-row records and weights (16 + 8 * sizeof(T) bytes per target row) and the row
    // load order (4 bytes per scheduled row) in shared memory
    // (the two-phase kernel keeps only the weights its instance can use: 4 per row for bands of <= 4 row taps)
    const int w_per_row = (panel_cta && rows.k_max <= 4) ? 4 : kRingMaxTaps;  // = RMAX of the instance picked below
    // (lanes instances of <= 5 row taps: 6 weights per row, 16-bit load order -- must match the kernel's carve-up)
    const bool lanes_small = lanes && (lanes2 || rows.k_max <= 5) && rows.n_src + 2 < 65536;
    const size_t row_tables =
        lanes_small ? (static_cast<size_t>(rows.n_out) * (16 + kLaneWrecSmall * sizeof(T)) +
                       (static_cast<size_t>(ring_c->n_sched) * sizeof(uint16_t) + 15) / 16 * 16)
                    : (static_cast<size_t>(rows.n_out) * (16 + (lanes ? kRingMaxTaps : w_per_row) * sizeof(T)) +
                       static_cast<size_t>(ring_c->n_sched) * sizeof(int32_t));
    // Several CTAs per SM = several independent row pipelines.  fp32 rows carry half the bytes of fp64 ones and
    // low coarsening ratios bring only two new source rows per target row: either way a target row leaves too
    // little time for ONE CTA's per-row dependency chain, so the shared memory is split between 2 - 3 CTAs when
    // a ring deep enough for the band (rows alive + 3 in flight) still fits
    int ctas = 1, slots = 0;
    for (int c = min_ctas; c >= 1; --c) {
      const long long budget = c > 1 ? (static_cast<long long>(dev.max_smem_sm) / c - 1024)
                                     : static_cast<long long>(dev.max_smem_optin);
      const long long room = budget - static_cast<long long>(probe.total) - static_cast<long long>(row_tables);
      int sl = room > 0 ? static_cast<int>(room / probe.row_stride) : 0;
      if (sl > kRingMaxSlots) sl = kRingMaxSlots;
      if (sl >= ring_c->max_live + (c > 1 ? 3 : 1) && sl >= 2) {
        ctas = c;
        slots = sl;
        break;
      }
    }
    if (slots >= ring_c->max_live + 1 && slots >= 2) {
      const RingLayout L =
          lanes ? ring_layout<T>(wp, 31, false, 0, 0, slots)
                : ring_layout<T>(wp, ring_c->pad_shift, skipna != 0, ring_c->strip_cols_max, ring_c->n_warps, slots);
      static const int lane_split = [] {  // (A/B switch of the split latitude pass: RG_LANE_SPLIT=0|1|2 points)
        const char* e = std::getenv("RG_LANE_SPLIT");
        return e ? std::atoi(e) : 2;
      }();
      // (A/B switch: RG_L2_HINT=0 drops the evict-first hint of the producer's bulk copies)
      static const int l2_hint = std::getenv("RG_L2_HINT") ? std::atoi(std::getenv("RG_L2_HINT")) : 1;
      RingDev ring{ring_c->sched_row,
                   reinterpret_cast<const int4*>(ring_c->krec),
                   ring_c->wrec,
                   ring_c->run_first,
                   ring_c->run_sched,
                   ring_c->strip_l0,
                   ring_c->strip_c0,
                   ring_c->strip_nc,
                   ring_c->n_runs,
                   ring_c->n_sched,
                   ring_c->max_live,
                   ring_c->pad_shift,
                   ring_c->n_strips,
                   ring_c->strip_cols_max,
                   ring_c->n_warps,
                   ring_c->lane_cols,
                   pole,
                   rows.n_src,
                   zero_fill,
                   ring_c->panel_strip0,
                   ring_c->panel_c0,
                   ring_c->panel_nc,
                   n_panels,
                   wp,
                   ring_c->strip_b0,
                   ring_c->block_l0,
                   ring_c->halo_l,
                   ring_c->halo_r,
                   lane_split,
                   l2_hint};
      if (!have_panels) return RG_ERR_NULL;  // (the host layer always provides the panel tables)
      using RingKernel = void (*)(const T*, T*, long long, long long, long long, long long, Band<T>, RingDev,
                                  T, int, int);
      const int warps = ring.n_warps + 2;  // + producer + sync
      const bool small = warps <= 14;
      RingKernel kern = nullptr;
      // strips of more than 32 targets (panel layouts only) need the two-pass longitude phase (TG = 2)
      const bool tg2 = panel_cta && ring_c->strip_targets_max > 32;
#define RG_PANEL(KW, RMAXV, MINBV, TGV)                                                               \
  (skipna ? conservative_ring_kernel<T, true, KW, 8, RMAXV, MINBV, TGV>                               \
          : conservative_ring_kernel<T, false, KW, 8, RMAXV, MINBV, TGV>)
#define RG_PICK(KW)                                                                                   \
  case KW:                                                                                            \
    if (panel_cta && rows.k_max <= 4 && KW <= 3)                                                      \
      kern = tg2 ? RG_PANEL(KW, 4, 3, 2) : RG_PANEL(KW, 4, 3, 1);                                     \
    else if (panel_cta && rows.k_max <= 4)                                                            \
      kern = tg2 ? RG_PANEL(KW, 4, 2, 2) : RG_PANEL(KW, 4, 3, 1);                                     \
    else if (panel_cta)                                                                               \
      kern = tg2 ? RG_PANEL(KW, 8, 2, 2) : RG_PANEL(KW, 8, 2, 1);                                     \
    else                                                                                              \
      kern = skipna ? (small ? conservative_ring_kernel<T, true, KW, 14, 8, 1, 1>                     \
                             : conservative_ring_kernel<T, true, KW, 26, 8, 1, 1>)                    \
                    : (small ? conservative_ring_kernel<T, false, KW, 14, 8, 1, 1>                    \
                             : conservative_ring_kernel<T, false, KW, 26, 8, 1, 1>);                  \
    break;
      switch (kw) {
        RG_PICK(1) RG_PICK(2) RG_PICK(3) RG_PICK(5) RG_PICK(8)
        default: break;
      }
#undef RG_PANEL
#undef RG_PICK
      const long long tasks = static_cast<long long>(batch) * ring.n_runs * n_panels;
      long long grid = static_cast<long long>(dev.sm_count) * ctas;
      if (grid > tasks) grid = tasks;
      if (lanes) {
        const size_t smem_l = L.total + row_tables;
        RG_CUDA(cudaFuncSetAttribute(lk, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_l)));
        if (lanes2 && n_panels > 1) {  // a CTA stays on one panel: task % n_panels must not depend on the iteration
          if (grid < n_panels) return RG_ERR_UNSUPPORTED;  // (tasks >= n_panels always)
          grid -= grid % n_panels;
        }
        lk<<<static_cast<unsigned>(grid), warps * 32, smem_l, stream>>>(x, y, batch, xbs, xrs, ybs, cols, ring, thr,
                                                                       slots, rows.n_out);
        return static_cast<int>(cudaGetLastError());
      }
      const size_t smem_r = L.total + row_tables;
      RG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_r)));
      kern<<<static_cast<unsigned>(grid), warps * 32, smem_r, stream>>>(x, y, batch, xbs, xrs, ybs, cols, ring, thr,
                                                                       slots, rows.n_out);
      return static_cast<int>(cudaGetLastError());
    }
  }
  if (smem > static_cast<size_t>(dev.max_smem_optin)) return RG_ERR_SMEM;

  using Kernel = void (*)(const T*, T*, long long, long long, long long, long long, Band<T>,
                          Band<T>, T, const T*, int, int);
  Kernel kern;
  const bool few = rows.k_max <= 4;
  if (aligned) {
    kern = skipna ? (few ? conservative_kernel<T, true, kMaxVec, 4> : conservative_kernel<T, true, kMaxVec, 8>)
                  : (few ? conservative_kernel<T, false, kMaxVec, 4> : conservative_kernel<T, false, kMaxVec, 8>);
  } else {
    kern = skipna ? (few ? conservative_kernel<T, true, 1, 4> : conservative_kernel<T, true, 1, 8>)
                  : (few ? conservative_kernel<T, false, 1, 4> : conservative_kernel<T, false, 1, 8>);
  }
  if (smem > 48 * 1024) {
    RG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(smem)));
  }
  int per_sm = 0;
  RG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kConsThreads, smem));
  if (per_sm < 1) return RG_ERR_SMEM;
  const long long tasks = static_cast<long long>(batch) * rows.n_out;
  long long grid = static_cast<long long>(dev.sm_count) * per_sm;
  if (grid > tasks) grid = tasks;
  kern<<<static_cast<unsigned>(grid), kConsThreads, smem, stream>>>(x, y, batch, xbs, xrs, ybs, rows,
                                                                   cols, thr, pole, w_pad, zero_fill);
  return static_cast<int>(cudaGetLastError());
}

// ------------------------------------------------------ host-buffer streaming pipeline
inline size_t round256(size_t n) { return (n + 255) / 256 * 256; }

// Chunked H2D -> (pole rows) -> kernel -> D2H over the three streams of an rg_pipeline.  `pipe` may be a
// caller-owned persistent pipeline (streams / events created once, the caller may defer the final drain with
// RG_HOST_NO_DRAIN) or NULL (a temporary one is created and destroyed by the call).
template <typename T>
int conservative_host(const T* xh, T* yh, int64_t batch, int64_t chunk, const rg_band_t* rows,
                      const rg_band_t* cols, int32_t skipna, T thr, int32_t pole_rows, int32_t south_row,
                      int32_t north_row, const rg_row_ring_t* ring, void* ws, size_t ws_bytes,
                      rg_pipeline* pipe, int32_t flags) {
  if (!xh || !yh || !ws || !rows || !cols) return RG_ERR_NULL;
  if (batch < 0 || chunk <= 0) return RG_ERR_SHAPE;
  if (batch == 0) return RG_OK;
  rg_pipeline local;
  if (!pipe) {
    const int rc = local.create(3);
    if (rc != RG_OK) { local.destroy(); return rc; }
    pipe = &local;
  }
  const int n_slots = pipe->n_slots;
  const int64_t H = rows->n_src, W = cols->n_src, Ho = rows->n_out, Wo = cols->n_out;
  const size_t in_b = round256(static_cast<size_t>(chunk) * H * W * sizeof(T));
  const size_t out_b = round256(static_cast<size_t>(chunk) * Ho * Wo * sizeof(T));
  const size_t pole_b = round256(static_cast<size_t>(chunk) * 2 * W * sizeof(T));
  int rc = RG_OK;
  if (ws_bytes < n_slots * (in_b + out_b + pole_b)) rc = RG_ERR_SHAPE;
  else if (reinterpret_cast<uintptr_t>(ws) % 256 != 0) rc = RG_ERR_ALIGN;

  unsigned char* base = static_cast<unsigned char*>(ws);
  const int64_t n_chunks = (batch + chunk - 1) / chunk;
  for (int64_t i = 0; i < n_chunks && rc == RG_OK; ++i) {
    const int s = static_cast<int>(i % n_slots);
    T* d_in = reinterpret_cast<T*>(base + s * (in_b + out_b + pole_b));
    T* d_out = reinterpret_cast<T*>(base + s * (in_b + out_b + pole_b) + in_b);
    T* d_pole = reinterpret_cast<T*>(base + s * (in_b + out_b + pole_b) + in_b + out_b);
    const int64_t b0 = i * chunk;
    const int64_t nb = (batch - b0 < chunk) ? batch - b0 : chunk;
    // ONE copy per chunk and direction (never the batched-copy entry points)
    rc = pipe->h2d(s, d_in, xh + b0 * H * W, static_cast<size_t>(nb) * H * W * sizeof(T));
    if (rc == RG_OK) rc = pipe->compute_begin(s);
    if (rc == RG_OK && pole_rows)
      rc = row_nanmean_launch<T>(d_in, d_pole, nb, static_cast<int32_t>(H), static_cast<int32_t>(W), H * W, W,
                                 south_row, north_row, pipe->s_k);
    if (rc == RG_OK)
      rc = conservative_launch<T>(d_in, d_out, nb, H * W, W, Ho * Wo, rows, cols,
                                  skipna ? kNanSkip : kNanZeroFill, thr, pole_rows ? d_pole : nullptr, ring,
                                  pipe->s_k);
    if (rc == RG_OK) rc = pipe->compute_end(s);
    if (rc == RG_OK) rc = pipe->d2h(s, yh + b0 * Ho * Wo, d_out, static_cast<size_t>(nb) * Ho * Wo * sizeof(T));
  }
  // after an error nothing may stay in flight into the caller's buffers
  if (rc != RG_OK || pipe == &local || !(flags & RG_HOST_NO_DRAIN)) {
    const int rd = pipe->drain();
    if (rc == RG_OK) rc = rd;
  }
  if (pipe == &local) local.destroy();
  return rc;
}

}  // namespace rg

// ------------------------------------------------------------------------------ C ABI
extern "C" {

int rg_conservative_f64(const double* x, double* y, int64_t batch, int64_t xbs, int64_t xrs,
                        int64_t ybs, const rg_band_t* rows, const rg_band_t* cols, int32_t skipna,
                        double valid_thr, const double* pole, const rg_row_ring_t* ring, void* stream) {
  return rg::conservative_launch<double>(x, y, batch, xbs, xrs, ybs, rows, cols,
                                         skipna ? rg::kNanSkip : rg::kNanZeroFill, valid_thr, pole, ring,
                                         static_cast<cudaStream_t>(stream));
}

int rg_conservative_f32(const float* x, float* y, int64_t batch, int64_t xbs, int64_t xrs,
                        int64_t ybs, const rg_band_t* rows, const rg_band_t* cols, int32_t skipna,
                        float valid_thr, const float* pole, const rg_row_ring_t* ring, void* stream) {
  return rg::conservative_launch<float>(x, y, batch, xbs, xrs, ybs, rows, cols,
                                        skipna ? rg::kNanSkip : rg::kNanZeroFill, valid_thr, pole, ring,
                                        static_cast<cudaStream_t>(stream));
}

int rg_row_nanmean_f64(const double* x, double* pole, int64_t batch, int32_t H, int32_t W,
                       int64_t xbs, int64_t xrs, int32_t south_row, int32_t north_row, void* stream) {
  return rg::row_nanmean_launch<double>(x, pole, batch, H, W, xbs, xrs, south_row, north_row,
                                        static_cast<cudaStream_t>(stream));
}

int rg_row_nanmean_f32(const float* x, float* pole, int64_t batch, int32_t H, int32_t W,
                       int64_t xbs, int64_t xrs, int32_t south_row, int32_t north_row, void* stream) {
  return rg::row_nanmean_launch<float>(x, pole, batch, H, W, xbs, xrs, south_row, north_row,
                                       static_cast<cudaStream_t>(stream));
}

int rg_separable_f64(const double* x, double* y, int64_t batch, int64_t xbs, int64_t xrs, int64_t ybs,
                     const rg_band_t* rows, const rg_band_t* cols, const double* pole,
                     const rg_row_ring_t* ring, void* stream) {
  return rg::conservative_launch<double>(x, y, batch, xbs, xrs, ybs, rows, cols, rg::kNanPropagate, 0.0,
                                         pole, ring, static_cast<cudaStream_t>(stream));
}

int rg_separable_f32(const float* x, float* y, int64_t batch, int64_t xbs, int64_t xrs, int64_t ybs,
                     const rg_band_t* rows, const rg_band_t* cols, const float* pole,
                     const rg_row_ring_t* ring, void* stream) {
  return rg::conservative_launch<float>(x, y, batch, xbs, xrs, ybs, rows, cols, rg::kNanPropagate, 0.0f,
                                        pole, ring, static_cast<cudaStream_t>(stream));
}

size_t rg_conservative_host_workspace(int64_t chunk_batch, int32_t H, int32_t W, int32_t Ho,
                                      int32_t Wo, int32_t elem_size, int32_t n_slots) {
  if (chunk_batch <= 0 || H <= 0 || W <= 0 || Ho <= 0 || Wo <= 0 || elem_size <= 0 || n_slots <= 0) return 0;
  const size_t in_b = rg::round256(static_cast<size_t>(chunk_batch) * H * W * elem_size);
  const size_t out_b = rg::round256(static_cast<size_t>(chunk_batch) * Ho * Wo * elem_size);
  const size_t pole_b = rg::round256(static_cast<size_t>(chunk_batch) * 2 * W * elem_size);
  return static_cast<size_t>(n_slots) * (in_b + out_b + pole_b);
}

int rg_conservative_host_f64(const double* x_host, double* y_host, int64_t batch,
                             int64_t chunk_batch, const rg_band_t* rows, const rg_band_t* cols,
                             int32_t skipna, double valid_thr, int32_t pole_rows, int32_t south_row,
                             int32_t north_row, const rg_row_ring_t* ring, void* workspace,
                             size_t workspace_bytes, rg_pipeline_t* pipeline, int32_t flags) {
  return rg::conservative_host<double>(x_host, y_host, batch, chunk_batch, rows, cols, skipna,
                                       valid_thr, pole_rows, south_row, north_row, ring, workspace,
                                       workspace_bytes, pipeline, flags);
}

int rg_conservative_host_f32(const float* x_host, float* y_host, int64_t batch,
                             int64_t chunk_batch, const rg_band_t* rows, const rg_band_t* cols,
                             int32_t skipna, float valid_thr, int32_t pole_rows, int32_t south_row,
                             int32_t north_row, const rg_row_ring_t* ring, void* workspace,
                             size_t workspace_bytes, rg_pipeline_t* pipeline, int32_t flags) {
  return rg::conservative_host<float>(x_host, y_host, batch, chunk_batch, rows, cols, skipna,
                                      valid_thr, pole_rows, south_row, north_row, ring, workspace,
                                      workspace_bytes, pipeline, flags);
}

}  // extern "C"
