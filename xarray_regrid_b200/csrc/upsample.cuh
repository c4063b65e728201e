// Separable small-band UPSAMPLING kernel (sm_100a): linear (2 x 2 taps) and cubic evaluation (4 x 4 taps)
// onto a much finer grid, e.g. 1 deg -> 0.1 deg (config 3: 181 x 360 -> 1801 x 3600).
//
// The op is write-bound (100 output bytes per input byte), so the kernel is organised around the stores:
//   * a CTA owns (slab b, band of R consecutive output rows) and ALL output columns;
//   * every thread owns G groups of 4 consecutive output columns for the whole band and keeps their
//     longitude taps (index + weight) in registers -- loaded once per CTA, the CTA is persistent;
//   * the few source rows the band needs are staged in shared memory (sliding window, coalesced loads);
//   * per output row: (A) one latitude-interpolated source row z[j] = sum_t wr[t] * xs[row_t][j] is
//     built in shared memory (double buffered -> one barrier per output row), (B) each thread produces
//     its outputs as sum_c wc[c] * z[idx[c]] and writes them with 16-byte streaming stores.
// NaN propagates through every tap (weight 0 included), which is the interpolation semantics
// (scipy interp1d: w_hi*y_hi + w_lo*y_lo).  Replaces data.interp(...) of methods/interp.py:43-46.
//
// Roofline: HBM (writes).  Algorithmic bytes per slab = (H*W + Ho*Wo) * sizeof(T).
#pragma once

#include "common.cuh"

namespace rg {

constexpr int kUpMaxThreads = 512;
constexpr int kUpRowCap = 16;      // source rows held in shared memory
constexpr int kUpBandRows = 32;    // output rows per task
constexpr int kUpRowsPerSync = 4;  // output rows produced per barrier (RPS; 8 for the 512-thread fp64 instances:
                                   // 8 x 181 (row, vector) pairs fill three rounds of 512 threads to 94 %, 4 x 181 two
                                   // rounds to 71 %, and the CTA meets at half as many barriers)

template <typename T> struct Vec4;
template <> struct alignas(16) Vec4<float> { float v[4]; };
template <> struct alignas(16) Vec4<double> { double v[4]; };

__device__ __forceinline__ void st_stream4(float* p, const Vec4<float>& a) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a.v[0]), "f"(a.v[1]), "f"(a.v[2]),
               "f"(a.v[3])
               : "memory");
}
__device__ __forceinline__ void st_stream4(double* p, const Vec4<double>& a) {
  if ((reinterpret_cast<uintptr_t>(p) & 31) == 0) {  // one 32-byte store (sm_100: STG.256) when the group is aligned
    asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a.v[0]), "d"(a.v[1]), "d"(a.v[2]),
                 "d"(a.v[3])
                 : "memory");
    return;
  }
  asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a.v[0]), "d"(a.v[1]) : "memory");
  asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p + 2), "d"(a.v[2]), "d"(a.v[3]) : "memory");
}

// Rows [kb, kb + nk) of a band: latitude pass into the z buffers of parity BUF, one barrier, longitude
// pass + stores.  BUF and the row slot are compile-time so that every z read is LDS [reg + immediate].
// WIN (1 or 2): the four outputs of a group take their taps from one window of KC + WIN consecutive z columns (host contract,
// rg_band_t.group_span): the window is loaded once per group (KC + WIN loads instead of 4 * KC) and every output is a
// (KC + WIN)-term dot product with its taps shifted into place (zeros elsewhere).  Used for the 4-tap cubic evaluation,
// where a NaN anywhere blanks the whole result anyway, so a zero weight times a NaN neighbour cannot matter.
template <typename T, int KR, int KC, int G, int ZW, int NT, int BUF, int WIN, int RPS>
__device__ __forceinline__ void up_rows(T* __restrict__ y_band, int kb_local, int nk, int W, int Wo,
                                        const T* s_x, T* s_z, const int (*s_off)[KR], const T (*s_rw)[KR],
                                        const int* s_state, const int (&ci)[G][4][KC],
                                        const T (&cw)[G][4][KC + WIN], int tid, T nan) {
  // ---- (A) latitude pass.  Only W values per row: 16-byte vectors, all rows of the step in flight at once
  // (the pass is latency bound, not throughput bound: few warps take part, the others go to the barrier)
  constexpr int V = 16 / sizeof(T);
  if (W % V == 0) {
    const int nvec = W / V;
    // (row, vector) pairs are dealt round-robin over ALL threads: the pass is short and sits in front of a
    // barrier, so its critical path (not its instruction count) is what the other warps wait for
    for (int p = tid; p < nvec * nk; p += NT) {
      int q = 0;
#pragma unroll
      for (int i = 1; i < RPS; ++i) q += (p >= i * nvec) ? 1 : 0;
      const int j = p - q * nvec;
      const int m = kb_local + q;
      Pack<T, V> acc;
#pragma unroll
      for (int e = 0; e < V; ++e) acc.v[e] = T(0);
#pragma unroll
      for (int t = 0; t < KR; ++t) {
        const T w = s_rw[m][t];
        const Pack<T, V> xv = *reinterpret_cast<const Pack<T, V>*>(s_x + s_off[m][t] + j * V);
#pragma unroll
        for (int e = 0; e < V; ++e) acc.v[e] = fma(w, xv.v[e], acc.v[e]);
      }
      *reinterpret_cast<Pack<T, V>*>(s_z + (BUF * RPS + q) * ZW + j * V) = acc;
    }
  } else {
    for (int q = 0; q < nk; ++q) {
      const int m = kb_local + q;
      int off[KR];
      T w[KR];
#pragma unroll
      for (int t = 0; t < KR; ++t) { off[t] = s_off[m][t]; w[t] = s_rw[m][t]; }
      T* z = s_z + (BUF * RPS + q) * ZW;
      for (int j = tid; j < W; j += NT) {
        T acc = T(0);
#pragma unroll
        for (int t = 0; t < KR; ++t) acc = fma(w[t], s_x[off[t] + j], acc);
        z[j] = acc;
      }
    }
  }
  __syncthreads();
  // ---- (B) longitude pass + 16-byte streaming stores
#pragma unroll
  for (int q = 0; q < RPS; ++q) {
    if (q < nk) {
      T* yrow = y_band + static_cast<long long>(kb_local + q) * Wo;
      const unsigned char* z = reinterpret_cast<const unsigned char*>(s_z + (BUF * RPS + q) * ZW);
      const int state = s_state[kb_local + q];
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int l = (tid + g * NT) * 4;
        if (l < Wo) {
          Vec4<T> out;
          if constexpr (WIN) {
            T zw[KC + WIN];
#pragma unroll
            for (int c = 0; c < KC + WIN; ++c)
              zw[c] = *reinterpret_cast<const T*>(z + ci[g][0][0] + c * static_cast<int>(sizeof(T)));
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              T acc = T(0);
#pragma unroll
              for (int c = 0; c < KC + WIN; ++c) acc = fma(cw[g][e][c], zw[c], acc);
              out.v[e] = acc;
            }
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              T acc = T(0);
#pragma unroll
              for (int c = 0; c < KC; ++c)
                acc = fma(cw[g][e][c], *reinterpret_cast<const T*>(z + ci[g][e][c]), acc);
              out.v[e] = acc;
            }
          }
          if (state != 2) {  // uniform: uncovered latitude -> NaN row; covered without taps -> 0 (or NaN column)
#pragma unroll
            for (int e = 0; e < 4; ++e) out.v[e] = (state == 0 || cw[g][e][0] != cw[g][e][0]) ? nan : T(0);
          }
          if (l + 4 <= Wo) {
            st_stream4(yrow + l, out);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)  // unrolled + predicated: keeps `out` in registers
              if (l + e < Wo) st_stream(yrow + l + e, out.v[e]);
          }
        }
      }
    }
  }
}

template <typename T, int KR, int KC, int G, int ZW, int NT, int WIN = 0, int RPS = kUpRowsPerSync>
__global__ void __launch_bounds__(NT)
upsample_kernel(const T* __restrict__ x, T* __restrict__ y, long long batch, long long xbs, long long xrs,
                long long ybs, Band<T> rows, Band<T> cols, int band_rows) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int H = rows.n_src, W = cols.n_src, Ho = rows.n_out, Wo = cols.n_out;
  T* s_z = reinterpret_cast<T*>(smem_raw);             // [2][RPS][ZW] latitude-interpolated rows
  T* s_x = s_z + 2 * RPS * ZW;              // [2][kUpRowCap][W] staged source rows (this / next band)
  // per-band row metadata, filled by warp 0 (double-buffered like the rows)
  __shared__ int s_off[2][kUpBandRows][KR];  // element offset of tap t's row inside the window
  __shared__ T s_rw[2][kUpBandRows][KR];
  __shared__ int s_state[2][kUpBandRows];    // 0: row not covered (NaN), 1: covered without taps (0), 2: normal
  __shared__ int s_win0[2];

  const int tid = threadIdx.x;
  const T nan = quiet_nan<T>();
  if constexpr (WIN) {  // z columns W, W + 1 of every z row: read by windows at the right edge, weight 0
    for (int i = tid; i < 2 * RPS * 2; i += NT) s_z[(i >> 1) * ZW + W + (i & 1)] = T(0);
  }

  // ---- longitude taps of the columns this thread owns (registers, loaded once): byte offsets into a
  //      z row and weights; a column outside the source range gets a NaN weight (NaN * z = NaN)
  int ci[G][4][KC];
  T cw[G][4][KC + WIN];
  if constexpr (WIN) {
    // window form: ci[g][0][0] = byte offset of the group's window, cw[g][e][c] = weight of window column c
#pragma unroll
    for (int g = 0; g < G; ++g) {
      int first[4], n[4];
      bool covered[4];
      int base = W;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int l = (tid + g * NT) * 4 + e;
        const bool in = l < Wo;
        n[e] = in ? min(cols.cnt[l], KC) : 0;
        covered[e] = in && cols.cover[l] != 0;
        first[e] = n[e] > 0 ? cols.idx[l] : W;
        if (covered[e]) base = min(base, first[e]);
      }
      if (base >= W) base = 0;  // (the window may reach 2 columns past W: those z columns are kept at zero)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int l = (tid + g * NT) * 4 + e;
        const int d = first[e] - base;  // host contract: 0 <= d, d + taps <= KC + WIN for covered outputs
#pragma unroll
        for (int c = 0; c < KC + WIN; ++c) {
          const int t = c - d;
          cw[g][e][c] = (covered[e] && t >= 0 && t < n[e]) ? cols.w[t * Wo + l] : T(0);
        }
        if (l < Wo && !covered[e]) cw[g][e][0] = nan;
#pragma unroll
        for (int c = 0; c < KC; ++c) ci[g][e][c] = 0;
      }
      ci[g][0][0] = base * static_cast<int>(sizeof(T));
    }
  } else {
#pragma unroll
    for (int g = 0; g < G; ++g) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int l = (tid + g * NT) * 4 + e;
        const bool in = l < Wo;
        const int n = in ? cols.cnt[l] : 0;
#pragma unroll
        for (int c = 0; c < KC; ++c) {
          const bool live = c < n && c < cols.k_max;
          ci[g][e][c] = (live ? cols.idx[c * Wo + l] : (n > 0 ? cols.idx[l] : 0)) * static_cast<int>(sizeof(T));
          cw[g][e][c] = live ? cols.w[c * Wo + l] : T(0);
        }
        if (in && !cols.cover[l]) cw[g][e][0] = nan;
      }
    }
  }

  // ---- band loop.  The row taps and the staged source rows of the NEXT band are prefetched while the current
  // band is produced (tables: warp 0 issues the loads at the band start and parks them in registers until its
  // first step is done; rows: LDGSTS into the other window buffer), so that no global-load latency sits between
  // two bands.  Bands with fewer than two steps fall back to the synchronous prologue.
  const int n_bands = (Ho + band_rows - 1) / band_rows;
  const long long tasks = batch * n_bands;
  const int x_elems = kUpRowCap * W;  // one window buffer

  // row taps of lane `tid`'s row of a band -> registers (warp 0 only); rmin reduced over the warp
  auto meta_load = [&](long long task, int (&ri)[KR], T (&rw)[KR], int& n, bool& cov) {
    const int band = static_cast<int>(task % n_bands);
    const int k0 = band * band_rows, k1 = min(k0 + band_rows, Ho);
    const int k = k0 + tid;
    n = 0;
    cov = false;
#pragma unroll
    for (int t = 0; t < KR; ++t) { ri[t] = 0; rw[t] = T(0); }
    if (k < k1) {
      cov = rows.cover[k] != 0;
      n = rows.cnt[k];
#pragma unroll
      for (int t = 0; t < KR; ++t) {
        if (t < rows.k_max) {
          ri[t] = rows.idx[t * Ho + k];
          rw[t] = rows.w[t * Ho + k];
        }
      }
    }
  };
  auto meta_store = [&](long long task, int buf, const int (&ri_in)[KR], const T (&rw_in)[KR], int n_in, bool cov) {
    const int band = static_cast<int>(task % n_bands);
    const int k0 = band * band_rows, k1 = min(k0 + band_rows, Ho);
    const int k = k0 + tid;
    const int n = cov ? n_in : 0;
    int ri[KR];
    T rw[KR];
    int rmin = H;
#pragma unroll
    for (int t = 0; t < KR; ++t) {
      const bool live = t < n && t < rows.k_max;
      ri[t] = live ? ri_in[t] : (n > 0 ? ri_in[0] : 0);
      rw[t] = live ? rw_in[t] : T(0);
      if (k < k1 && n > 0) rmin = min(rmin, ri[t]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rmin = min(rmin, __shfl_xor_sync(0xffffffffu, rmin, o));
    if (rmin >= H) rmin = 0;
    if (k < k1) {
#pragma unroll
      for (int t = 0; t < KR; ++t) {
        // host contract (window_rows): < kUpRowCap rows; clamped so a broken table cannot leave the window
        s_off[buf][tid][t] = (n > 0 ? min(ri[t] - rmin, kUpRowCap - 1) : 0) * W;
        s_rw[buf][tid][t] = rw[t];
      }
      s_state[buf][tid] = !cov ? 0 : (n == 0 ? 1 : 2);
    }
    if (tid == 0) s_win0[buf] = rmin;
  };
  // the window's source rows -> s_x[buf], asynchronously (one element per request: any alignment)
  // (the source is small and re-read by every band while gigabytes of output stream through L2: evict-last)
  unsigned long long keep;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep));
  auto rows_issue = [&](long long task, int buf) {
    const T* xb = x + (task / n_bands) * xbs;
    const int w0 = s_win0[buf];
    const int wn = min(kUpRowCap, H - w0);
    T* dst = s_x + buf * x_elems;
    for (int r = 0; r < wn; ++r) {
      const T* src = xb + static_cast<long long>(w0 + r) * xrs;
      for (int j = tid; j < W; j += NT) {
        const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst + r * W + j));
        asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], %2, %3;" ::"r"(d), "l"(src + j),
                     "n"(sizeof(T)), "l"(keep)
                     : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int cur = 0;
  bool ready = false;  // is the current band's prologue already in buffer `cur`?
  for (long long task = blockIdx.x; task < tasks; task += gridDim.x) {
    const int band = static_cast<int>(task % n_bands);
    const long long b = task / n_bands;
    const int k0 = band * band_rows, k1 = min(k0 + band_rows, Ho);
    const long long next = task + gridDim.x;
    const bool prefetch = next < tasks && (k1 - k0) > RPS;  // needs two steps in this band

    if (!ready) {  // synchronous prologue (first band of the CTA, or after a one-step band)
      if (tid < 32) {
        int ri[KR];
        T rw[KR];
        int n;
        bool cov;
        meta_load(task, ri, rw, n, cov);
        meta_store(task, cur, ri, rw, n, cov);
      }
      __syncthreads();
      rows_issue(task, cur);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    // next band's row taps: loads issued now (warp 0), consumed after the first step
    int nri[KR];
    T nrw[KR];
    int nn = 0;
    bool ncov = false;
    if (prefetch && tid < 32) meta_load(next, nri, nrw, nn, ncov);

    T* y_band = y + b * ybs + static_cast<long long>(k0) * Wo;
    const T* sx = s_x + cur * x_elems;
    for (int kl = 0; kl < k1 - k0; kl += 2 * RPS) {
      up_rows<T, KR, KC, G, ZW, NT, 0, WIN, RPS>(y_band, kl, min(RPS, k1 - k0 - kl), W, Wo, sx, s_z, s_off[cur],
                                            s_rw[cur], s_state[cur], ci, cw, tid, nan);
      if (kl == 0 && prefetch && tid < 32) meta_store(next, cur ^ 1, nri, nrw, nn, ncov);
      const int kl2 = kl + RPS;
      if (kl2 < k1 - k0)
        up_rows<T, KR, KC, G, ZW, NT, 1, WIN, RPS>(y_band, kl2, min(RPS, k1 - k0 - kl2), W, Wo, sx, s_z,
                                              s_off[cur], s_rw[cur], s_state[cur], ci, cw, tid, nan);
      // the barrier inside the second step ordered warp 0's table writes: everybody can start the row copies
      if (kl == 0 && prefetch) rows_issue(next, cur ^ 1);
    }
    __syncthreads();  // all reads of this band's window and z rows are done
    ready = prefetch;
    if (prefetch) cur ^= 1;
  }
}

// Eligibility + launch.  Returns RG_ERR_UNSUPPORTED when the generic kernels should be used instead.
template <typename T>
int upsample_try_launch(const T* x, T* y, int64_t batch, int64_t xbs, int64_t xrs, int64_t ybs,
                        const Band<T>& rows, const Band<T>& cols, const DeviceInfo& dev,
                        cudaStream_t stream) {
  const int W = cols.n_src, Wo = cols.n_out;
  if (rows.k_max > 4 || cols.k_max > 4) return RG_ERR_UNSUPPORTED;
  // the host guarantees (rg_band_t.window_rows) that every aligned group of `band_rows` output rows takes its
  // taps from at most kUpRowCap consecutive source rows: the staged window
  int band_rows = rows.window_rows;
  if (band_rows < kUpRowsPerSync || band_rows % kUpRowsPerSync != 0) return RG_ERR_UNSUPPORTED;
  if (band_rows > kUpBandRows) band_rows = kUpBandRows;
  // heavy taps (4 fp64 taps per column) use 512-thread CTAs so that the per-thread tap registers halve
  const int nt = (cols.k_max >= 2 && sizeof(T) == 8) ? 512 : 256;
  if (Wo < 2 * W || Wo > nt * 4 * 4) return RG_ERR_UNSUPPORTED;  // upsampling, <= 4 groups / thread
  if ((reinterpret_cast<uintptr_t>(y) % 16) != 0 || (ybs * sizeof(T)) % 16 != 0 || (Wo * sizeof(T)) % 16 != 0)
    return RG_ERR_UNSUPPORTED;
  if (W > 1024) return RG_ERR_UNSUPPORTED;
  const int zw = W <= 512 ? 512 : 1024;
  size_t smem = (static_cast<size_t>(2 * kUpRowCap) * W + 2 * kUpRowsPerSync * zw) * sizeof(T);
  if (smem > static_cast<size_t>(dev.max_smem_optin)) return RG_ERR_UNSUPPORTED;
  // eight rows per barrier (512-thread instances, rows of <= 512 columns, bands of a multiple of 8 rows)
  const size_t smem8 = (static_cast<size_t>(2 * kUpRowCap) * W + 2 * 8 * zw) * sizeof(T);
  const bool rps8 = nt == 512 && zw == 512 && band_rows % 8 == 0 && band_rows >= 16 && smem8 <= static_cast<size_t>(dev.max_smem_optin);
  if (rps8) smem = smem8;
  const int groups = ((Wo + 3) / 4 + nt - 1) / nt;  // 1..4
  const int kr = rows.k_max <= 1 ? 1 : rows.k_max <= 2 ? 2 : 4;
  const int kc = cols.k_max <= 1 ? 1 : cols.k_max <= 2 ? 2 : 4;
  using Kernel = void (*)(const T*, T*, long long, long long, long long, long long, Band<T>, Band<T>, int);
  Kernel kern = nullptr;
  // 4 taps: window form of the longitude pass when every group of 4 outputs fits a 6-column window
  const bool win = kc == 4 && kr == 4 && cols.group_span > 0 && cols.group_span <= 6 && W + 2 <= zw;
  if (win) {
    if (nt == 512 && groups <= 2) {
      // (a 5-column window when the groups fit one -- x10 upsampling does: 20 instead of 24 weights and FMAs per group)
      const bool w5 = cols.group_span <= 5;
      if (rps8 && w5) kern = groups == 1 ? upsample_kernel<T, 4, 4, 1, 512, 512, 1, 8> : upsample_kernel<T, 4, 4, 2, 512, 512, 1, 8>;
      else if (rps8) kern = groups == 1 ? upsample_kernel<T, 4, 4, 1, 512, 512, 2, 8> : upsample_kernel<T, 4, 4, 2, 512, 512, 2, 8>;
      else if (groups == 1) kern = zw == 512 ? upsample_kernel<T, 4, 4, 1, 512, 512, 2> : upsample_kernel<T, 4, 4, 1, 1024, 512, 2>;
      else kern = zw == 512 ? upsample_kernel<T, 4, 4, 2, 512, 512, 2> : upsample_kernel<T, 4, 4, 2, 1024, 512, 2>;
    }
  }
#define RG_UP(KR_, KC_, G_)                                                                              \
  if (!kern && kr == KR_ && kc == KC_ && groups == G_) {                                                          \
    if (nt == 256)                                                                                       \
      kern = zw == 512 ? upsample_kernel<T, KR_, KC_, G_, 512, 256> : upsample_kernel<T, KR_, KC_, G_, 1024, 256>; \
    else if (G_ <= 2 && rps8)                                                                            \
      kern = upsample_kernel<T, KR_, KC_, (G_ <= 2 ? G_ : 1), 512, 512, 0, 8>;                           \
    else if (G_ <= 2)                                                                                    \
      kern = zw == 512 ? upsample_kernel<T, KR_, KC_, (G_ <= 2 ? G_ : 1), 512, 512>                      \
                       : upsample_kernel<T, KR_, KC_, (G_ <= 2 ? G_ : 1), 1024, 512>;                    \
  }
#define RG_UP_G(KR_, KC_) RG_UP(KR_, KC_, 1) RG_UP(KR_, KC_, 2) RG_UP(KR_, KC_, 3) RG_UP(KR_, KC_, 4)
  RG_UP_G(1, 1) RG_UP_G(2, 2) RG_UP_G(4, 4)
#undef RG_UP_G
#undef RG_UP
  if (!kern) return RG_ERR_UNSUPPORTED;  // mixed tap counts: generic kernels
  if (smem > 48 * 1024)
    RG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  int per_sm = 0;
  RG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, nt, smem));
  if (per_sm < 1) return RG_ERR_UNSUPPORTED;
  const int n_bands = (rows.n_out + band_rows - 1) / band_rows;
  const long long tasks = static_cast<long long>(batch) * n_bands;
  long long grid = static_cast<long long>(dev.sm_count) * per_sm;
  if (grid > tasks) grid = tasks;
  kern<<<static_cast<unsigned>(grid), nt, smem, stream>>>(x, y, batch, xbs, xrs, ybs, rows, cols,
                                                                 band_rows);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace rg
