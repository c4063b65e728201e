// Shared device / host helpers for the regrid_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "regrid_b200.h"

#define RG_CUDA(call)                          \
  do {                                         \
    cudaError_t rg_e_ = (call);                \
    if (rg_e_ != cudaSuccess) return (int)rg_e_; \
  } while (0)

namespace rg {

// ---------------------------------------------------------------- device info (cached)
struct DeviceInfo {
  int sm_count;
  int max_smem_optin;  // per CTA
  int max_smem_sm;     // per SM
};
int device_info(DeviceInfo* out);  // defined in api.cu

// Typed view of rg_band_t.
template <typename T>
struct Band {
  const int32_t* __restrict__ idx;
  const T* __restrict__ w;
  const int32_t* __restrict__ cnt;
  const uint8_t* __restrict__ cover;
  int n_out, k_max, n_src, span, window_rows, group_span;
};

template <typename T>
inline int make_band(const rg_band_t* b, Band<T>* out) {
  if (!b || !b->idx || !b->w || !b->cnt || !b->cover) return RG_ERR_NULL;
  if (b->n_out <= 0 || b->k_max <= 0 || b->n_src <= 0) return RG_ERR_SHAPE;
  out->idx = b->idx;
  out->w = static_cast<const T*>(b->w);
  out->cnt = b->cnt;
  out->cover = b->cover;
  out->n_out = b->n_out;
  out->k_max = b->k_max;
  out->n_src = b->n_src;
  out->span = b->span;
  out->window_rows = b->window_rows;
  out->group_span = b->group_span;
  return RG_OK;
}

// --------------------------------------------------------------------- device helpers
template <typename T> __device__ __forceinline__ T quiet_nan();
template <> __device__ __forceinline__ double quiet_nan<double>() { return CUDART_NAN; }
template <> __device__ __forceinline__ float quiet_nan<float>() { return CUDART_NAN_F; }

// Streaming vector loads: read-only path, do not allocate in L1 (every source element is
// consumed once per CTA; reuse between neighbouring CTAs is served by the 126 MB L2).
template <typename T, int VEC> struct alignas(sizeof(T) * VEC) Pack { T v[VEC]; };

// L2 eviction policy of data that is read once (the compiler hoists the createpolicy out of loops)
__device__ __forceinline__ unsigned long long l2_evict_first() {
  unsigned long long pol;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

template <typename T, int VEC>
__device__ __forceinline__ Pack<T, VEC> ld_stream(const T* p);

template <> __device__ __forceinline__ Pack<double, 1> ld_stream<double, 1>(const double* p) {
  Pack<double, 1> r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(r.v[0]) : "l"(p), "l"(l2_evict_first()));
  return r;
}
template <> __device__ __forceinline__ Pack<double, 2> ld_stream<double, 2>(const double* p) {
  Pack<double, 2> r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;"
               : "=d"(r.v[0]), "=d"(r.v[1]) : "l"(p), "l"(l2_evict_first()));
  return r;
}
template <> __device__ __forceinline__ Pack<float, 1> ld_stream<float, 1>(const float* p) {
  Pack<float, 1> r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(r.v[0]) : "l"(p), "l"(l2_evict_first()));
  return r;
}
template <> __device__ __forceinline__ Pack<float, 2> ld_stream<float, 2>(const float* p) {
  Pack<float, 2> r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;"
               : "=f"(r.v[0]), "=f"(r.v[1]) : "l"(p), "l"(l2_evict_first()));
  return r;
}
template <> __device__ __forceinline__ Pack<float, 4> ld_stream<float, 4>(const float* p) {
  Pack<float, 4> r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]) : "l"(p), "l"(l2_evict_first()));
  return r;
}

// Streaming (evict-first) scalar store for outputs that are written once.
__device__ __forceinline__ void st_stream(double* p, double v) {
  asm volatile("st.global.cs.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void st_stream(float* p, float v) {
  asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void st_stream2(double* p, double a, double b) {
  asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void st_stream2(float* p, float a, float b) {
  asm volatile("st.global.cs.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}

}  // namespace rg
