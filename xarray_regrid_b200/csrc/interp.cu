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
