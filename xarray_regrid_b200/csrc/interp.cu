// Interpolation-side kernels for sm_100a: nearest-neighbour gather and the cubic spline's
// "any NaN -> all NaN" rule.  Linear and cubic themselves are separable banded contractions and run
// on the conservative kernels (rg_separable_* forwards to the same launcher with skipna = 0):
//   linear : 2-tap bands  (scipy interp1d._call_linear, reached from methods/interp.py:43-46)
//   cubic  : prefilter band (A^-1 of the not-a-knot collocation system) then 4-tap B-spline bands
//   nearest: this file -- a pure index gather, bit-exact, any integer or float element type.
//
// Roofline: HBM.  Algorithmic bytes per output cell = sizeof(Tin) + sizeof(Tout); the source is read
// with a stride of n_src / n_out elements, so every touched 32-byte sector is partly wasted when
// down-sampling (the "sector-touched" bound in SURVEY.md 8d).

#include "common.cuh"

namespace rg {

constexpr int kNearThreads = 256;

template <typename Tin, typename Tout>
__global__ void __launch_bounds__(kNearThreads)
nearest_kernel(const Tin* __restrict__ x, Tout* __restrict__ y, long long batch, long long xbs,
               long long xrs, long long ybs, int Ho, int Wo, const int32_t* __restrict__ row_idx,
               const uint8_t* __restrict__ row_ok, const int32_t* __restrict__ col_idx,
               const uint8_t* __restrict__ col_ok, Tout fill) {
  const long long tasks = batch * Ho;
  for (long long task = blockIdx.x; task < tasks; task += gridDim.x) {
    const int k = static_cast<int>(task % Ho);
    const long long b = task / Ho;
    Tout* yrow = y + b * ybs + static_cast<long long>(k) * Wo;
    if (!row_ok[k]) {  // CTA-uniform
      for (int l = threadIdx.x; l < Wo; l += kNearThreads) yrow[l] = fill;
      continue;
    }
    const Tin* xrow = x + b * xbs + static_cast<long long>(row_idx[k]) * xrs;
    for (int l = threadIdx.x; l < Wo; l += kNearThreads) {
      Tout v = fill;
      if (col_ok[l]) v = static_cast<Tout>(__ldg(xrow + col_idx[l]));
      yrow[l] = v;
    }
  }
}

template <typename Tin, typename Tout>
int nearest_launch(const void* x, void* y, int64_t batch, int64_t xbs, int64_t xrs, int64_t ybs,
                   int32_t Ho, int32_t Wo, const int32_t* row_idx, const uint8_t* row_ok,
                   const int32_t* col_idx, const uint8_t* col_ok, double fill, cudaStream_t stream) {
  DeviceInfo dev;
  int rc = device_info(&dev);
  if (rc != RG_OK) return rc;
  const long long tasks = static_cast<long long>(batch) * Ho;
  long long grid = static_cast<long long>(dev.sm_count) * 8;
  if (grid > tasks) grid = tasks;
  nearest_kernel<Tin, Tout><<<static_cast<unsigned>(grid), kNearThreads, 0, stream>>>(
      static_cast<const Tin*>(x), static_cast<Tout*>(y), batch, xbs, xrs, ybs, Ho, Wo, row_idx, row_ok,
      col_idx, col_ok, static_cast<Tout>(fill));
  return static_cast<int>(cudaGetLastError());
}

// ------------------------------------------------------------------------- NaN flag / poison
template <typename T>
__global__ void __launch_bounds__(256)
nan_flag_kernel(const T* __restrict__ x, long long batch, int H, int W, long long xbs, long long xrs,
                int32_t* __restrict__ flag) {
  const long long rows = batch * H;
  bool bad = false;
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    const T* row = x + (r / H) * xbs + (r % H) * xrs;
    for (int j = threadIdx.x; j < W; j += 256) {
      const T v = row[j];
      bad |= (v != v);
    }
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(flag, 1);
}

template <typename T>
__global__ void __launch_bounds__(256)
poison_kernel(T* __restrict__ y, long long n, const int32_t* __restrict__ flag) {
  if (*flag == 0) return;
  const T nan = quiet_nan<T>();
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += 256LL * gridDim.x) y[i] = nan;
}

// ------------------------------------------------------------------ spline coefficient solve
// The prefilter of the cubic method: scipy's make_interp_spline solves the banded not-a-knot collocation
// system A c = y (two sub- and two super-diagonals) for every line along the interpolated axis.  The host
// factors A = L U once per axis (no pivoting: the system is diagonally dominant; the host checks the
// factorisation against scipy), the device runs the forward / backward substitution: one thread per line,
// the five factor diagonals in shared memory (broadcast reads), values prefetched 8 at a time so the
// recurrence is the only serial dependency (one FMA per element on the critical path: the term of the older
// value is folded in first).  O(n) per line instead of the O(n * 63) of the truncated inverse.
//   forward   z_i = y_i - l1_i z_{i-1} - l2_i z_{i-2}
//   backward  c_i = (z_i - u1_i c_{i+1} - u2_i c_{i+2}) * dinv_i
// Factors as the kernels keep them in shared memory: [5][n] = -l1, -l2, dinv, -u1 * dinv, -u2 * dinv, so that a step
// of either substitution has ONE fused multiply-add on its critical path:
//   forward   z_i = fma(-l1_i, z_{i-1}, fma(-l2_i, z_{i-2}, y_i))
//   backward  c_i = fma(-u1_i dinv_i, c_{i+1}, fma(-u2_i dinv_i, c_{i+2}, z_i dinv_i))
__device__ __forceinline__ void spline_stage_factors(double* s_fac, const double* __restrict__ lower,
                                                     const double* __restrict__ upper, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double d = upper[i];
    s_fac[i] = -lower[i];
    s_fac[n + i] = -lower[n + i];
    s_fac[2 * n + i] = d;
    s_fac[3 * n + i] = -upper[n + i] * d;
    s_fac[4 * n + i] = -upper[2 * n + i] * d;
  }
}

// One substitution over the elements [lo, hi) of a line p[0], p[step], ... (forward: ascending from the state
// (s1, s2) = (z_{lo-1}, z_{lo-2}); backward: descending from (c_hi, c_{hi+1})).  Batches of 8: the next batch's values
// and factors are fetched before the current batch's dependent chain runs (the solve is in place, so the compiler
// may not move loads above the stores by itself); full batches carry no bounds test on the chain.  STORE = false
// only advances the state (warm-up over a neighbouring segment, see the tile kernel).
template <bool STORE, typename Ptr>
__device__ __forceinline__ void spline_forward(Ptr p, long long step, int lo, int hi, const double* l1,
                                               const double* l2, double& z1, double& z2) {
  double v[8], a1[8], a2[8], vn[8], a1n[8], a2n[8];
  const int full = lo + ((hi - lo) & ~7);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int i = min(lo + e, hi - 1);
    v[e] = p[i * step];
    a1[e] = l1[i];
    a2[e] = l2[i];
  }
  for (int i0 = lo; i0 < full; i0 += 8) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int i = min(i0 + 8 + e, hi - 1);
      vn[e] = p[i * step];
      a1n[e] = l1[i];
      a2n[e] = l2[i];
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const double zi = fma(a1[e], z1, fma(a2[e], z2, v[e]));  // the inner term does not wait for z1
      z2 = z1;
      z1 = zi;
      if constexpr (STORE) p[(i0 + e) * step] = zi;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      v[e] = vn[e];
      a1[e] = a1n[e];
      a2[e] = a2n[e];
    }
  }
  for (int i = full; i < hi; ++i) {
    const double zi = fma(l1[i], z1, fma(l2[i], z2, p[i * step]));
    z2 = z1;
    z1 = zi;
    if constexpr (STORE) p[i * step] = zi;
  }
}

template <bool STORE, typename Ptr>
__device__ __forceinline__ void spline_backward(Ptr p, long long step, int lo, int hi, const double* dinv,
                                                const double* u1, const double* u2, double& c1, double& c2) {
  double v[8], a1[8], a2[8], d[8], vn[8], a1n[8], a2n[8], dn[8];
  const int tail = (hi - lo) & 7;  // the descending walk starts with the partial batch
  for (int i = hi - 1; i >= hi - tail; --i) {
    const double ci = fma(u1[i], c1, fma(u2[i], c2, p[i * step] * dinv[i]));
    c2 = c1;
    c1 = ci;
    if constexpr (STORE) p[i * step] = ci;
  }
  const int top = hi - tail - 1;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int i = max(top - e, lo);
    v[e] = p[i * step];
    a1[e] = u1[i];
    a2[e] = u2[i];
    d[e] = dinv[i];
  }
  for (int i0 = top; i0 >= lo; i0 -= 8) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int i = max(i0 - 8 - e, lo);
      vn[e] = p[i * step];
      a1n[e] = u1[i];
      a2n[e] = u2[i];
      dn[e] = dinv[i];
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const double ci = fma(a1[e], c1, fma(a2[e], c2, v[e] * d[e]));
      c2 = c1;
      c1 = ci;
      if constexpr (STORE) p[(i0 - e) * step] = ci;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      v[e] = vn[e];
      a1[e] = a1n[e];
      a2[e] = a2n[e];
      d[e] = dn[e];
    }
  }
}

template <typename Ptr>
__device__ __forceinline__ void spline_solve_line(Ptr p, long long step, int n, const double* l1, const double* l2,
                                                  const double* dinv, const double* u1, const double* u2) {
  double s1 = 0.0, s2 = 0.0;
  spline_forward<true>(p, step, 0, n, l1, l2, s1, s2);
  s1 = 0.0;
  s2 = 0.0;
  spline_backward<true>(p, step, 0, n, dinv, u1, u2, s1, s2);
}

template <int AXIS>
__global__ void __launch_bounds__(128)
spline_solve_kernel(double* __restrict__ y, long long batch, long long ybs, long long yrs, int H, int W,
                    const double* __restrict__ lower, const double* __restrict__ upper) {
  extern __shared__ double s_fac[];  // [5][n], see spline_stage_factors
  const int n = AXIS == 0 ? H : W;
  const int per_slab = AXIS == 0 ? W : H;
  spline_stage_factors(s_fac, lower, upper, n);
  __syncthreads();
  const double *l1 = s_fac, *l2 = s_fac + n, *dinv = s_fac + 2 * n, *u1 = s_fac + 3 * n, *u2 = s_fac + 4 * n;
  const long long line = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (line >= batch * per_slab) return;
  const long long b = line / per_slab;
  const int j = static_cast<int>(line % per_slab);
  double* p = y + b * ybs + (AXIS == 0 ? static_cast<long long>(j) : static_cast<long long>(j) * yrs);
  spline_solve_line(p, AXIS == 0 ? yrs : 1, n, l1, l2, dinv, u1, u2);
}

// Lines along the contiguous axis: a CTA stages kTileLines whole lines in shared memory with coalesced
// loads (odd row pitch: the per-line walkers hit distinct banks), runs both substitutions out of shared memory and
// writes the tile back.  One read and one write of the data.
//
// Four threads per line (warp s = segment s of every line of the tile): both recurrences are contractions (interior
// factors of a spline on a near-regular axis: 0.268 per step), so a segment may start from a ZERO state kWarm
// elements early -- what it forgets is below 0.35^40 = 6e-19 of the data -- instead of waiting for the segment before
// it.  Phases, a CTA barrier between them: forward warm-up over the previous segment's tail (read only), forward
// over the own segment (in place), backward warm-up over the next segment's head (read only), backward over the own
// segment.  262 instead of 724 dependent steps per thread for 362-element lines (ncu before: one warp per CTA
// walking the lines, 4.3 cycles per instruction of its serial stream, 66 us).  The CTA checks the factors of every
// warm-up window first and walks whole lines with one thread each when a window is not contractive enough
// (irregular axes) or the line is too short.
constexpr int kTileLines = 32;
constexpr int kSplineWarm = 40;
constexpr double kSplineRho = 0.35;  // max |a1| + |a2| per step inside a warm-up window
__global__ void __launch_bounds__(128)
spline_solve_tile_kernel(double* __restrict__ y, long long n_lines, int H, long long ybs, long long yrs, int W,
                         const double* __restrict__ lower, const double* __restrict__ upper) {
  extern __shared__ double s_fac[];  // [5][W] factors, then [kTileLines][W | 1] data
  const int n = W, pitch = W | 1;
  double* tile = s_fac + 5 * n;
  spline_stage_factors(s_fac, lower, upper, n);
  const long long line0 = static_cast<long long>(blockIdx.x) * kTileLines;
  const int lines = static_cast<int>(min(static_cast<long long>(kTileLines), n_lines - line0));
  // all lines of the tile in flight at once (LDGSTS, 8 bytes per request: the odd pitch allows no wider one); with
  // plain loads + stores the lines were fetched one after the other and the kernel spent its time on their latency
  // (ncu: long-scoreboard 9.5 warps per issue, 11 % issue activity)
  // (slab, row) of the tile's lines advance with a carry: a 64-bit division per line was most of the kernel's
  // instructions (ncu: 34 K warp instructions per tile, 8 K of them per warp in these two copy loops)
  const long long b0 = line0 / H;
  const int r0 = static_cast<int>(line0 - b0 * H);
  {
    const double* src = y + b0 * ybs + static_cast<long long>(r0) * yrs;
    int r = r0;
    const uint32_t t0 = static_cast<uint32_t>(__cvta_generic_to_shared(tile));
    for (int l = 0; l < lines; ++l) {
      for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const uint32_t d = t0 + static_cast<uint32_t>(l * pitch + j) * 8u;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src + j) : "memory");
      }
      src += yrs;
      if (++r == H) {  // next slab
        r = 0;
        src += ybs - static_cast<long long>(H) * yrs;
      }
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  const double *l1 = s_fac, *l2 = s_fac + n, *dinv = s_fac + 2 * n, *u1 = s_fac + 3 * n, *u2 = s_fac + 4 * n;
  // segments [s * seg, (s + 1) * seg): are all warm-up windows inside the line and contractive?
  const int seg = (n + 3) / 4;
  int bad = (seg < kSplineWarm + 1 || 3 * seg + kSplineWarm > n) ? 1 : 0;
  if (!bad) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int r = i % seg, s = i / seg;
      const bool fwd_win = s <= 2 && r >= seg - kSplineWarm;  // tail of segments 0..2: forward warm-up of s + 1
      const bool bwd_win = s >= 1 && r < kSplineWarm;         // head of segments 1..3: backward warm-up of s - 1
      if (fwd_win && fabs(l1[i]) + fabs(l2[i]) > kSplineRho) bad = 1;
      if (bwd_win && fabs(u1[i]) + fabs(u2[i]) > kSplineRho) bad = 1;
    }
  }
  bad = __syncthreads_or(bad);
  if (bad) {
    if (threadIdx.x < lines) spline_solve_line(tile + threadIdx.x * pitch, 1, n, l1, l2, dinv, u1, u2);
    __syncthreads();
  } else {
    const int s = threadIdx.x >> 5, l = threadIdx.x & 31;
    const bool on = l < lines;
    double* p = tile + (on ? l : 0) * pitch;
    const int lo = s * seg, hi = min(lo + seg, n);
    double s1 = 0.0, s2 = 0.0;
    if (on && s > 0) spline_forward<false>(p, 1, lo - kSplineWarm, lo, l1, l2, s1, s2);
    __syncthreads();
    if (on) spline_forward<true>(p, 1, lo, hi, l1, l2, s1, s2);
    __syncthreads();
    s1 = 0.0;
    s2 = 0.0;
    if (on && s < 3) spline_backward<false>(p, 1, hi, hi + kSplineWarm, dinv, u1, u2, s1, s2);
    __syncthreads();
    if (on) spline_backward<true>(p, 1, lo, hi, dinv, u1, u2, s1, s2);
    __syncthreads();
  }
  {
    double* dst = y + b0 * ybs + static_cast<long long>(r0) * yrs;
    int r = r0;
    for (int l = 0; l < lines; ++l) {
      for (int j = threadIdx.x; j < n; j += blockDim.x) dst[j] = tile[l * pitch + j];
      dst += yrs;
      if (++r == H) {
        r = 0;
        dst += ybs - static_cast<long long>(H) * yrs;
      }
    }
  }
}

}  // namespace rg

extern "C" {

int rg_nearest(const void* x, void* y, int32_t in_type, int32_t out_type, int64_t batch,
               int64_t x_batch_stride, int64_t x_row_stride, int64_t y_batch_stride, int32_t Ho,
               int32_t Wo, const int32_t* row_idx, const uint8_t* row_ok, const int32_t* col_idx,
               const uint8_t* col_ok, double fill, void* stream) {
  if (!x || !y || !row_idx || !row_ok || !col_idx || !col_ok) return RG_ERR_NULL;
  if (batch < 0 || Ho <= 0 || Wo <= 0 || x_row_stride <= 0 ||
      y_batch_stride < static_cast<int64_t>(Ho) * Wo)
    return RG_ERR_SHAPE;
  if (batch == 0) return RG_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (out_type == in_type) {
    switch (in_type) {
      case RG_U8: return rg::nearest_launch<uint8_t, uint8_t>(x, y, batch, x_batch_stride, x_row_stride, y_batch_stride, Ho, Wo, row_idx, row_ok, col_idx, col_ok, fill, s);
      case RG_I8: return rg::nearest_launch<int8_t, int8_t>(x, y, batch, x_batch_stride, x_row_stride, y_batch_stride, Ho, Wo, row_idx, row_ok, col_idx, col_ok, fill, s);
      case RG_I16: return rg::nearest_launch<int16_t, int16_t>(x, y, batch, x_batch_stride, x_row_stride, y_batch_stride, Ho, Wo, row_idx, row_ok, col_idx, col_ok, fill, s);
      case RG_I32: return rg::nearest_launch<int32_t, int32_t>(x, y, batch, x_batch_stride, x_row_stride, y_batch_stride, Ho, Wo, row_idx, row_ok, col_idx, col_ok, fill, s);
      case RG_I64: return rg::nearest_launch<long long, long long>(x, y, batch, x_batch_stride, x_row_stride, y_batch_stride, Ho, Wo, row_idx, row_ok, col_idx, col_ok, fill, s);
      case RG_F32: return rg::nearest_launch<float, float>(x, y, batch, x_batch_stride, x_row_stride, y_batch_stride, Ho, Wo, row_idx, row_ok, col_idx, col_ok, fill, s);
      case RG_F64: return rg::nearest_launch<double, double>(x, y, batch, x_batch_stride, x_row_stride, y_batch_stride, Ho, Wo, row_idx, row_ok, col_idx, col_ok, fill, s);
      default: return RG_ERR_UNSUPPORTED;
    }
  }
  if (out_type == RG_F64) {  // the reference's dtype rule: non-float input comes back float64
    switch (in_type) {
      case RG_U8: return rg::nearest_launch<uint8_t, double>(x, y, batch, x_batch_stride, x_row_stride, y_batch_stride, Ho, Wo, row_idx, row_ok, col_idx, col_ok, fill, s);
      case RG_I8: return rg::nearest_launch<int8_t, double>(x, y, batch, x_batch_stride, x_row_stride, y_batch_stride, Ho, Wo, row_idx, row_ok, col_idx, col_ok, fill, s);
      case RG_I16: return rg::nearest_launch<int16_t, double>(x, y, batch, x_batch_stride, x_row_stride, y_batch_stride, Ho, Wo, row_idx, row_ok, col_idx, col_ok, fill, s);
      case RG_I32: return rg::nearest_launch<int32_t, double>(x, y, batch, x_batch_stride, x_row_stride, y_batch_stride, Ho, Wo, row_idx, row_ok, col_idx, col_ok, fill, s);
      case RG_I64: return rg::nearest_launch<long long, double>(x, y, batch, x_batch_stride, x_row_stride, y_batch_stride, Ho, Wo, row_idx, row_ok, col_idx, col_ok, fill, s);
      case RG_F32: return rg::nearest_launch<float, double>(x, y, batch, x_batch_stride, x_row_stride, y_batch_stride, Ho, Wo, row_idx, row_ok, col_idx, col_ok, fill, s);
      default: return RG_ERR_UNSUPPORTED;
    }
  }
  return RG_ERR_UNSUPPORTED;
}

int rg_spline_solve_f64(double* y, int64_t batch, int64_t ybs, int64_t yrs, int32_t H, int32_t W, int32_t axis,
                        const double* lower, const double* upper, void* stream) {
  if (!y || !lower || !upper) return RG_ERR_NULL;
  if (batch < 0 || H <= 0 || W <= 0 || (axis != 0 && axis != 1) || yrs < W) return RG_ERR_SHAPE;
  if (batch == 0) return RG_OK;
  const int n = axis == 0 ? H : W;
  const size_t smem = static_cast<size_t>(5) * n * sizeof(double);
  rg::DeviceInfo dev;
  const int rc = rg::device_info(&dev);
  if (rc != RG_OK) return rc;
  if (smem > static_cast<size_t>(dev.max_smem_optin)) return RG_ERR_UNSUPPORTED;  // caller uses the band pass
  if (axis == 1) {  // contiguous lines: the staged-tile kernel when a tile fits shared memory
    const size_t smem_t = smem + static_cast<size_t>(rg::kTileLines) * (W | 1) * sizeof(double);
    if (smem_t <= static_cast<size_t>(dev.max_smem_optin)) {
      if (smem_t > 48 * 1024)
        RG_CUDA(cudaFuncSetAttribute(rg::spline_solve_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(smem_t)));
      const long long n_lines = static_cast<long long>(batch) * H;
      const long long tiles = (n_lines + rg::kTileLines - 1) / rg::kTileLines;
      if (tiles > 0x7fffffffLL) return RG_ERR_SHAPE;
      rg::spline_solve_tile_kernel<<<static_cast<unsigned>(tiles), 128, smem_t, static_cast<cudaStream_t>(stream)>>>(
          y, n_lines, H, ybs, yrs, W, lower, upper);
      return static_cast<int>(cudaGetLastError());
    }
  }
  auto kern = axis == 0 ? rg::spline_solve_kernel<0> : rg::spline_solve_kernel<1>;
  if (smem > 48 * 1024)
    RG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const long long lines = static_cast<long long>(batch) * (axis == 0 ? W : H);
  const long long blocks = (lines + 127) / 128;
  if (blocks > 0x7fffffffLL) return RG_ERR_SHAPE;
  kern<<<static_cast<unsigned>(blocks), 128, smem, static_cast<cudaStream_t>(stream)>>>(y, batch, ybs, yrs, H, W,
                                                                                     lower, upper);
  return static_cast<int>(cudaGetLastError());
}

int rg_nan_flag_f64(const double* x, int64_t batch, int32_t H, int32_t W, int64_t xbs, int64_t xrs,
                    int32_t* flag, void* stream) {
  if (!x || !flag) return RG_ERR_NULL;
  if (batch < 0 || H <= 0 || W <= 0) return RG_ERR_SHAPE;
  if (batch == 0) return RG_OK;
  long long rows = static_cast<long long>(batch) * H;
  const unsigned grid = static_cast<unsigned>(rows < 148 * 8 ? rows : 148 * 8);
  rg::nan_flag_kernel<double><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, batch, H, W, xbs, xrs, flag);
  return static_cast<int>(cudaGetLastError());
}

int rg_nan_flag_f32(const float* x, int64_t batch, int32_t H, int32_t W, int64_t xbs, int64_t xrs,
                    int32_t* flag, void* stream) {
  if (!x || !flag) return RG_ERR_NULL;
  if (batch < 0 || H <= 0 || W <= 0) return RG_ERR_SHAPE;
  if (batch == 0) return RG_OK;
  long long rows = static_cast<long long>(batch) * H;
  const unsigned grid = static_cast<unsigned>(rows < 148 * 8 ? rows : 148 * 8);
  rg::nan_flag_kernel<float><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, batch, H, W, xbs, xrs, flag);
  return static_cast<int>(cudaGetLastError());
}

int rg_poison_if_f64(double* y, int64_t n, const int32_t* flag, void* stream) {
  if (!y || !flag) return RG_ERR_NULL;
  if (n <= 0) return RG_OK;
  const long long blocks = (n + 255) / 256;
  rg::poison_kernel<double><<<static_cast<unsigned>(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0,
                              static_cast<cudaStream_t>(stream)>>>(y, n, flag);
  return static_cast<int>(cudaGetLastError());
}

int rg_poison_if_f32(float* y, int64_t n, const int32_t* flag, void* stream) {
  if (!y || !flag) return RG_ERR_NULL;
  if (n <= 0) return RG_OK;
  const long long blocks = (n + 255) / 256;
  rg::poison_kernel<float><<<static_cast<unsigned>(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0,
                             static_cast<cudaStream_t>(stream)>>>(y, n, flag);
  return static_cast<int>(cudaGetLastError());
}

}  // extern "C"
