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
template <int AXIS>
__global__ void __launch_bounds__(128)
spline_solve_kernel(double* __restrict__ y, long long batch, long long ybs, long long yrs, int H, int W,
                    const double* __restrict__ lower, const double* __restrict__ upper) {
  extern __shared__ double s_fac[];  // [5][n]: l1, l2, dinv, u1, u2
  const int n = AXIS == 0 ? H : W;
  const int per_slab = AXIS == 0 ? W : H;
  for (int i = threadIdx.x; i < 2 * n; i += blockDim.x) s_fac[i] = lower[i];
  for (int i = threadIdx.x; i < 3 * n; i += blockDim.x) s_fac[2 * n + i] = upper[i];
  __syncthreads();
  const double *l1 = s_fac, *l2 = s_fac + n, *dinv = s_fac + 2 * n, *u1 = s_fac + 3 * n, *u2 = s_fac + 4 * n;
  const long long line = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (line >= batch * per_slab) return;
  const long long b = line / per_slab;
  const int j = static_cast<int>(line % per_slab);
  double* p = y + b * ybs + (AXIS == 0 ? static_cast<long long>(j) : static_cast<long long>(j) * yrs);
  const long long step = AXIS == 0 ? yrs : 1;

  // The solve is in place, so the compiler may not move the next 8 loads above this batch's stores: they are issued
  // by hand one batch ahead (the batches touch disjoint elements), otherwise every batch waits a full DRAM latency.
  double z1 = 0.0, z2 = 0.0;
  double v[8], vn[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = (e < n) ? p[e * step] : 0.0;
  for (int i0 = 0; i0 < n; i0 += 8) {
#pragma unroll
    for (int e = 0; e < 8; ++e) vn[e] = (i0 + 8 + e < n) ? p[(i0 + 8 + e) * step] : 0.0;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int i = i0 + e;
      if (i < n) {
        const double zi = fma(-l1[i], z1, fma(-l2[i], z2, v[e]));  // the inner term does not wait for z1
        z2 = z1;
        z1 = zi;
        p[i * step] = zi;
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = vn[e];
  }
  double c1 = 0.0, c2 = 0.0;
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = (n - 1 - e >= 0) ? p[(n - 1 - e) * step] : 0.0;
  for (int i0 = n - 1; i0 >= 0; i0 -= 8) {
#pragma unroll
    for (int e = 0; e < 8; ++e) vn[e] = (i0 - 8 - e >= 0) ? p[(i0 - 8 - e) * step] : 0.0;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int i = i0 - e;
      if (i >= 0) {
        const double ci = fma(-u1[i], c1, fma(-u2[i], c2, v[e])) * dinv[i];
        c2 = c1;
        c1 = ci;
        p[i * step] = ci;
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = vn[e];
  }
}

// Lines along the contiguous axis: a CTA stages kTileLines whole lines in shared memory with coalesced
// loads (odd row pitch: the per-line walkers hit distinct banks), one thread per line runs both
// substitutions out of shared memory, the CTA writes the tile back.  One read and one write of the data.
constexpr int kTileLines = 32;
__global__ void __launch_bounds__(128)
spline_solve_tile_kernel(double* __restrict__ y, long long n_lines, int H, long long ybs, long long yrs, int W,
                         const double* __restrict__ lower, const double* __restrict__ upper) {
  extern __shared__ double s_fac[];  // [5][W] factors, then [kTileLines][W | 1] data
  const int n = W, pitch = W | 1;
  double* tile = s_fac + 5 * n;
  for (int i = threadIdx.x; i < 2 * n; i += blockDim.x) s_fac[i] = lower[i];
  for (int i = threadIdx.x; i < 3 * n; i += blockDim.x) s_fac[2 * n + i] = upper[i];
  const long long line0 = static_cast<long long>(blockIdx.x) * kTileLines;
  const int lines = static_cast<int>(min(static_cast<long long>(kTileLines), n_lines - line0));
  // all lines of the tile in flight at once (LDGSTS, 8 bytes per request: the odd pitch allows no wider one); with
  // plain loads + stores the lines were fetched one after the other and the kernel spent its time on their latency
  // (ncu: long-scoreboard 9.5 warps per issue, 11 % issue activity)
  for (int l = 0; l < lines; ++l) {
    const long long line = line0 + l;
    const double* src = y + (line / H) * ybs + (line % H) * yrs;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
      const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(tile + l * pitch + j));
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src + j) : "memory");
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < lines) {
    const double *l1 = s_fac, *l2 = s_fac + n, *dinv = s_fac + 2 * n, *u1 = s_fac + 3 * n, *u2 = s_fac + 4 * n;
    double* p = tile + threadIdx.x * pitch;
    // values and factors are fetched 8 at a time so that only the recurrence itself is serial
    double z1 = 0.0, z2 = 0.0;
    for (int i0 = 0; i0 < n; i0 += 8) {
      double v[8], a1[8], a2[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int i = min(i0 + e, n - 1);
        v[e] = p[i];
        a1[e] = l1[i];
        a2[e] = l2[i];
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const double zi = fma(-a1[e], z1, fma(-a2[e], z2, v[e]));  // the inner term does not wait for z1
        if (i0 + e < n) {
          z2 = z1;
          z1 = zi;
          p[i0 + e] = zi;
        }
      }
    }
    double c1 = 0.0, c2 = 0.0;
    for (int i0 = n - 1; i0 >= 0; i0 -= 8) {
      double v[8], a1[8], a2[8], d[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int i = max(i0 - e, 0);
        v[e] = p[i];
        a1[e] = u1[i];
        a2[e] = u2[i];
        d[e] = dinv[i];
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const double ci = fma(-a1[e], c1, fma(-a2[e], c2, v[e])) * d[e];
        if (i0 - e >= 0) {
          c2 = c1;
          c1 = ci;
          p[i0 - e] = ci;
        }
      }
    }
  }
  __syncthreads();
  for (int l = 0; l < lines; ++l) {
    const long long line = line0 + l;
    double* dst = y + (line / H) * ybs + (line % H) * yrs;
    for (int j = threadIdx.x; j < n; j += blockDim.x) dst[j] = tile[l * pitch + j];
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
