// Group-by regridders for sm_100a: most_common / least_common (label mode) and zonal statistics
// (sum, mean, var, std, min, max, median) per target cell.
//
// Replaces flox.xarray.xarray_reduce(... expected_groups=<closed-left bins> ...) as called from
// src/xarray_regrid/methods/flox_reduce.py:89-96 (statistics) and :174-183 (count per label, then
// idxmax / idxmin).  Bin membership is decided on the HOST with the reference's own expressions
// (np.digitize on `append(t, t[-1]+step) - step/2`, _shared.py:12-18), so a target cell is simply a
// contiguous range of (sorted) source rows x a contiguous range of (sorted, possibly wrapped) source
// columns -- the kernels never compare coordinates.
//
// One CTA stages the tile [rows of target row k] x [columns of a chunk of target cells] in shared
// memory with coalesced loads, then one warp per target cell
//   mode:   maps each element to its dense label index, finds equal labels inside the warp with
//           match.any (no atomics: one elected lane per distinct label adds the popcount to a warp-
//           private counter array), then arg-max / arg-min with ties to the first label;
//   stats:  shuffle reductions (two-pass variance, exactly like np.var), bitonic sort for the median.
//
// Roofline: HBM.  Algorithmic bytes = source cells * sizeof(T) + target cells * sizeof(Tout).

#include "common.cuh"

namespace rg {

struct Bins {
  const int32_t* __restrict__ start;  // [n_out] first formatted source position of the bin
  const int32_t* __restrict__ count;  // [n_out] number of positions
  const uint8_t* __restrict__ cover;  // [n_out]
  const int32_t* __restrict__ out_index;  // [n_out] where the bin goes in the output
  int n_out, n_src, src_offset, src_step;  // source index = (offset + step * position) mod n_src
};

inline int make_bins(const rg_bins_t* b, Bins* out) {
  if (!b || !b->start || !b->count || !b->cover || !b->out_index) return RG_ERR_NULL;
  if (b->n_out <= 0 || b->n_src <= 0 || (b->src_step != 1 && b->src_step != -1)) return RG_ERR_SHAPE;
  *out = Bins{b->start, b->count, b->cover, b->out_index, b->n_out, b->n_src, b->src_offset, b->src_step};
  return RG_OK;
}

__device__ __forceinline__ int src_index(const Bins& a, int pos) {
  int i = a.src_offset + a.src_step * pos;
  i %= a.n_src;
  return i < 0 ? i + a.n_src : i;
}

constexpr int kRedThreads = 256;
constexpr int kRedWarps = kRedThreads / 32;

enum { OP_SUM = 0, OP_MEAN = 1, OP_VAR = 2, OP_STD = 3, OP_MIN = 4, OP_MAX = 5, OP_MEDIAN = 6 };

// ------------------------------------------------------------------ tile staging (all ops)
// tile[r * ncol + c] = x[b, src_row(r0 + r), src_col(p0 + c)]
template <typename T>
__device__ __forceinline__ void load_tile(T* tile, const T* __restrict__ xb, long long xrs, const Bins& rows,
                                          const Bins& cols, int r0, int nr, int p0, int ncol) {
  for (int r = threadIdx.x >> 5; r < nr; r += kRedWarps) {
    const T* row = xb + static_cast<long long>(src_index(rows, r0 + r)) * xrs;
    T* trow = tile + r * ncol;
    for (int c = threadIdx.x & 31; c < ncol; c += 32) trow[c] = __ldg(row + src_index(cols, p0 + c));
  }
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------- mode
// label -> dense index in `values`: identity with bound (dense 0..n-1 labels), else binary search.
template <typename T>
__device__ __forceinline__ int label_index(T v, const T* __restrict__ values, int n_values, int dense) {
  if (dense) return (v >= T(0) && static_cast<long long>(v) < n_values) ? static_cast<int>(v) : -1;
  int lo = 0, hi = n_values - 1;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1;
    const T m = values[mid];
    if (m == v) return mid;
    if (m < v) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}

template <typename Tin, typename Tout>
__global__ void __launch_bounds__(kRedThreads)
mode_kernel(const Tin* __restrict__ x, Tout* __restrict__ y, long long batch, long long xbs, long long xrs,
            long long ybs, Bins rows, Bins cols, int cols_per_chunk, int n_chunks,
            const Tin* __restrict__ values_sorted, const int32_t* __restrict__ value_rank,
            const Tin* __restrict__ values, int n_values, int dense, int anti, Tout fill, int tile_elems) {
  extern __shared__ __align__(16) unsigned char smem[];
  Tin* tile = reinterpret_cast<Tin*>(smem);
  int32_t* counts = reinterpret_cast<int32_t*>(smem + (static_cast<size_t>(tile_elems) * sizeof(Tin) + 15) / 16 * 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int32_t* cnt = counts + warp * n_values;
  const long long tasks = batch * rows.n_out * n_chunks;
  const int Wo = cols.n_out;

  for (long long task = blockIdx.x; task < tasks; task += gridDim.x) {
    const int chunk = static_cast<int>(task % n_chunks);
    const int k = static_cast<int>((task / n_chunks) % rows.n_out);
    const long long b = task / (static_cast<long long>(n_chunks) * rows.n_out);
    const int l0 = chunk * cols_per_chunk;
    const int l1 = min(l0 + cols_per_chunk, Wo);
    Tout* yrow = y + b * ybs + static_cast<long long>(rows.out_index[k]) * Wo;
    const bool row_ok = rows.cover[k] != 0;
    const int r0 = rows.start[k], nr = row_ok ? rows.count[k] : 0;
    const int p0 = cols.start[l0];
    const int ncol = cols.start[l1 - 1] + cols.count[l1 - 1] - p0;
    const bool fits = static_cast<long long>(nr) * ncol <= tile_elems;  // host contract; never overrun
    if (fits && nr > 0 && ncol > 0) load_tile(tile, x + b * xbs, xrs, rows, cols, r0, nr, p0, ncol);
    __syncthreads();

    for (int l = l0 + warp; l < l1; l += kRedWarps) {
      Tout out = fill;
      if (fits && row_ok && cols.cover[l]) {
        for (int i = lane; i < n_values; i += 32) cnt[i] = 0;
        __syncwarp();
        const int c0 = cols.start[l] - p0, nc = cols.count[l];
        for (int r = 0; r < nr; ++r) {
          const Tin* trow = tile + r * ncol + c0;
          for (int cb = 0; cb < nc; cb += 32) {
            const int c = cb + lane;
            int d = -1;
            if (c < nc) {
              d = label_index(trow[c], values_sorted, n_values, dense);
              if (d >= 0 && !dense) d = value_rank[d];  // position in the caller's `values` order
            }
            const unsigned active = __ballot_sync(0xffffffffu, d >= 0);
            if (d >= 0) {
              const unsigned same = __match_any_sync(active, d);
              if ((__ffs(same) - 1) == lane) cnt[d] += __popc(same);  // one elected lane per label
            }
            __syncwarp();
          }
        }
        // arg-max (arg-min) over labels; empty labels count as -1 (flox fill_value=-1), ties -> first
        int best_i = 0x7fffffff;
        long long best_key = anti ? 0x7fffffffffffffffLL : -0x7fffffffffffffffLL;
        for (int i = lane; i < n_values; i += 32) {
          const int cval = cnt[i] > 0 ? cnt[i] : -1;
          const bool better = anti ? (cval < best_key) : (cval > best_key);
          if (better) { best_key = cval; best_i = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const long long ok = __shfl_xor_sync(0xffffffffu, best_key, o);
          const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
          const bool better = anti ? (ok < best_key) : (ok > best_key);
          if (better || (ok == best_key && oi < best_i)) { best_key = ok; best_i = oi; }
        }
        out = static_cast<Tout>(values[best_i]);
        __syncwarp();
      }
      if (lane == 0) yrow[cols.out_index[l]] = out;
    }
    __syncthreads();
  }
}

// Stage one source row segment of a strip (columns p0 .. p0 + ncol of the formatted axis) into shared
// memory.  Contiguous segments keep the source's 16-byte phase: element c lives at dst[shift + c] (shift is
// returned), so every aligned 16-byte chunk is copied with one LDG.128 + one STS.128.
template <typename T>
__device__ __forceinline__ int stage_row(T* dst, const T* __restrict__ src, const Bins& cols, int p0, int s0,
                                         int ncol, bool contiguous, int lane) {
  if (contiguous) {
    const T* seg = src + s0;
    constexpr int V = 16 / sizeof(T);
    const int shift = static_cast<int>((reinterpret_cast<uintptr_t>(seg) & 15) / sizeof(T));
    const int4* gsrc = reinterpret_cast<const int4*>(seg - shift);
    int4* sdst = reinterpret_cast<int4*>(dst);
    const int nvec = (shift + ncol + V - 1) / V;  // may read a few elements before / after the segment
    for (int v = lane; v < nvec; v += 32) sdst[v] = __ldg(gsrc + v);
    return shift;
  }
  for (int c = lane; c < ncol; c += 32) dst[c] = __ldg(src + src_index(cols, p0 + c));
  return 0;
}

// Asynchronous variant: contiguous segments go through LDGSTS (cp.async, 16 bytes per lane and request) so a
// warp can put several rows in flight without holding them in registers; the caller commits the group and
// waits (cp_async_wait) before reading.  Gathered segments fall back to the synchronous copy.
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  // (source cells are read exactly once: evict-first in L2)
  uint64_t pol;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <typename T>
__device__ __forceinline__ void stage_row_async(T* dst, const T* __restrict__ src, const Bins& cols, int p0, int s0,
                                                int ncol, bool contiguous, int lane) {
  if (contiguous) {
    const T* seg = src + s0;
    constexpr int V = 16 / sizeof(T);
    const int shift = static_cast<int>((reinterpret_cast<uintptr_t>(seg) & 15) / sizeof(T));
    const int4* gsrc = reinterpret_cast<const int4*>(seg - shift);
    int4* sdst = reinterpret_cast<int4*>(dst);
    const int nvec = (shift + ncol + V - 1) / V;
    for (int v = lane; v < nvec; v += 32) cp_async_16(sdst + v, gsrc + v);
    return;
  }
  for (int c = lane; c < ncol; c += 32) dst[c] = __ldg(src + src_index(cols, p0 + c));
}

// ---------------------------------------------------------------- mode, one cell per lane
// Throughput kernel for label counts: a warp owns a strip of 32 adjacent target cells of one target
// row, ONE CELL PER LANE.  Each lane keeps its own counters in shared memory laid out [label][thread]
// (bank = lane: conflict-free; the increment is a fire-and-forget shared atomic on a word no other lane
// touches, which is cheaper than LDS + IADD + STS; no voting); the warp stages kStageRows source rows of its strip at
// a time with 16-byte coalesced loads and every lane walks its own cell.  Needs n_values <= kLaneMaxLabels.
constexpr int kLaneThreads = 256;
constexpr int kLaneWarps = kLaneThreads / 32;
constexpr int kLaneMaxLabels = 64;
constexpr int kStageRows = 5;

template <typename Tin, typename Tout>
__global__ void __launch_bounds__(kLaneThreads)
mode_lanes_kernel(const Tin* __restrict__ x, Tout* __restrict__ y, long long batch, long long xbs, long long xrs,
                  long long ybs, Bins rows, Bins cols, const Tin* __restrict__ values_sorted,
                  const int32_t* __restrict__ value_rank, const Tin* __restrict__ values, int n_values, int dense,
                  int anti, Tout fill, int stage_elems) {
  extern __shared__ __align__(16) unsigned char smem[];
  uint32_t* s_cnt = reinterpret_cast<uint32_t*>(smem);  // [n_values][kLaneThreads]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  Tin* stage = reinterpret_cast<Tin*>(smem + static_cast<size_t>(n_values + 1) * kLaneThreads * 4) +
               static_cast<size_t>(warp) * kStageRows * stage_elems;
  const int Wo = cols.n_out;
  const int strips = (Wo + 31) / 32;
  const long long tasks = batch * rows.n_out * strips;  // one task = one 32-cell strip, handled by one warp
  const long long warp_id = static_cast<long long>(blockIdx.x) * kLaneWarps + warp;
  const long long n_warps = static_cast<long long>(gridDim.x) * kLaneWarps;

  for (long long task = warp_id; task < tasks; task += n_warps) {
    const int strip = static_cast<int>(task % strips);
    const int k = static_cast<int>((task / strips) % rows.n_out);
    const long long b = task / (static_cast<long long>(strips) * rows.n_out);
    const int l = strip * 32 + lane;
    const bool in = l < Wo;
    const int l_last = min(strip * 32 + 31, Wo - 1);
    const bool row_ok = rows.cover[k] != 0;
    const int r0 = rows.start[k], nr = row_ok ? rows.count[k] : 0;
    const int p0 = cols.start[strip * 32];
    const int ncol = cols.start[l_last] + cols.count[l_last] - p0;  // host contract: <= stage_elems
    const int c0 = in ? cols.start[l] - p0 : 0;
    const int nc = (in && cols.cover[l]) ? cols.count[l] : 0;
    int nc_max = nc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nc_max = max(nc_max, __shfl_xor_sync(0xffffffffu, nc_max, o));

    for (int i = 0; i <= n_values; ++i) s_cnt[i * kLaneThreads + tid] = 0;  // + 1 trash row
    const Tin* xb = x + b * xbs;
    // the strip's source columns are contiguous (no wrap inside the strip) in the common case
    const int s0 = src_index(cols, p0);
    const bool contiguous = cols.src_step == 1 && s0 + ncol <= cols.n_src;

    // lanes without a live cell count everything into their trash slot (stride 0)
    const bool live = nc > 0;
    const int lane_stride = live ? kLaneThreads : 0;
    uint32_t* mine = s_cnt + (live ? 0 : n_values * kLaneThreads) + tid;
    const int nc_eff = live ? nc : nc_max;
    int nc_min = nc_eff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nc_min = min(nc_min, __shfl_xor_sync(0xffffffffu, nc_min, o));
    const int row_first = src_index(rows, r0);  // rows of a bin never wrap: plain +-1 steps from here

    for (int rb = 0; rb < nr; rb += kStageRows) {
      const int nrb = min(kStageRows, nr - rb);
      __syncwarp();
      // ---- stage nrb rows x ncol columns (coalesced; 16-byte loads when the segment is aligned)
      for (int r = 0; r < nrb; ++r) {
        const Tin* src = xb + static_cast<long long>(row_first + rows.src_step * (rb + r)) * xrs;
        stage_row<Tin>(stage + r * stage_elems, src, cols, p0, s0, ncol, contiguous, lane);
      }
      __syncwarp();
      // ---- every lane walks its own cell.  Branch-free: labels that are not listed go to a trash counter
      //      row (index n_values); each step is extract + clamp + LDS / IADD / STS on a lane-private counter
      for (int r = 0; r < nrb; ++r) {
        int shift = 0;
        if (contiguous) {
          const Tin* seg = xb + static_cast<long long>(row_first + rows.src_step * (rb + r)) * xrs + s0;
          shift = static_cast<int>((reinterpret_cast<uintptr_t>(seg) & 15) / sizeof(Tin));
        }
        const int pos = shift + c0;
        const Tin* trow = stage + r * stage_elems + pos;
        if (dense) {
          for (int c = 0; c < nc_min; ++c) {  // every lane's cell has this column: no per-lane predicate
            const unsigned v = static_cast<unsigned>(static_cast<long long>(trow[c]));
            atomicAdd(mine + min(v, static_cast<unsigned>(n_values)) * lane_stride, 1u);
          }
          for (int c = nc_min; c < nc_max; ++c) {  // ragged tail
            const unsigned v = static_cast<unsigned>(static_cast<long long>(trow[c < nc_eff ? c : 0]));
            const unsigned d = (c < nc_eff && v < static_cast<unsigned>(n_values)) ? v : static_cast<unsigned>(n_values);
            atomicAdd(mine + d * lane_stride, 1u);  // fire-and-forget ATOMS on a lane-private word
          }
        } else {
          for (int c = 0; c < nc_max; ++c) {
            int d = label_index(trow[c < nc_eff ? c : 0], values_sorted, n_values, 0);
            d = (c < nc_eff && d >= 0) ? value_rank[d] : n_values;
            atomicAdd(mine + d * lane_stride, 1u);  // fire-and-forget ATOMS on a lane-private word
          }
        }
      }
    }
    // ---- arg-max (arg-min): empty labels count as -1 (flox fill_value=-1), ties -> first in `values`
    Tout out = fill;
    if (row_ok && in && cols.cover[l]) {
      int best_i = 0;
      int best = anti ? 0x7fffffff : -0x7fffffff;
      for (int i = 0; i < n_values; ++i) {
        const int cval = s_cnt[i * kLaneThreads + tid] > 0 ? static_cast<int>(s_cnt[i * kLaneThreads + tid]) : -1;
        const bool better = anti ? (cval < best) : (cval > best);
        if (better) { best = cval; best_i = i; }
      }
      out = static_cast<Tout>(values[best_i]);
    }
    if (in) y[b * ybs + static_cast<long long>(rows.out_index[k]) * Wo + cols.out_index[l]] = out;
  }
}

// ---- warp-level selection of order statistics (median) -------------------------------------------
// Values are mapped to order-preserving unsigned keys; every lane holds up to kSelCap keys of the cell in
// registers and the warp determines the k-th smallest key bit by bit (MSB first) with one REDUX per bit.
constexpr int kSelCap = 32;  // keys per lane -> cells of up to 1024 valid values

__device__ __forceinline__ unsigned int to_key(float v) {
  const unsigned int u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_key(unsigned int k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
__device__ __forceinline__ unsigned long long to_key(double v) {
  const unsigned long long u = static_cast<unsigned long long>(__double_as_longlong(v));
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double from_key(unsigned long long k) {
  return __longlong_as_double(static_cast<long long>((k >> 63) ? (k & 0x7fffffffffffffffull) : ~k));
}

// k-th smallest (0-based) of the valid keys spread over the warp's registers (bit i of `valid` marks slot i).
template <typename K, int CAP = kSelCap>
__device__ __forceinline__ K warp_select(const K (&keys)[CAP], unsigned valid, int k) {
  constexpr int BITS = sizeof(K) * 8;
  K prefix = 0;
  int cand = __reduce_add_sync(0xffffffffu, __popc(valid));  // keys still matching the prefix
  for (int b = BITS - 1; b >= 0; --b) {
    // keys that agree with `prefix` on the bits above b and have bit b clear
    const K want = prefix >> b;  // bit b of `want` is 0
    int cnt = 0;
#pragma unroll
    for (int i = 0; i < CAP; ++i) cnt += (((valid >> i) & 1u) && (keys[i] >> b) == want) ? 1 : 0;
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (k >= cnt) {
      k -= cnt;
      cand -= cnt;
      prefix |= static_cast<K>(1) << b;
    } else {
      cand = cnt;
    }
    if (cand == 1 && b > 0) {
      // one key left (distinct values: after ~log2(n) bits): it is the answer, fetch it instead of
      // resolving the remaining bits one by one
      K found = 0;
#pragma unroll
      for (int i = 0; i < CAP; ++i)
        if (((valid >> i) & 1u) && (keys[i] >> b) == (prefix >> b)) found = keys[i];
      if constexpr (sizeof(K) == 8) {
        const unsigned lo = __reduce_or_sync(0xffffffffu, static_cast<unsigned>(found));
        const unsigned hi = __reduce_or_sync(0xffffffffu, static_cast<unsigned>(found >> 32));
        return (static_cast<K>(hi) << 32) | lo;
      } else {
        return static_cast<K>(__reduce_or_sync(0xffffffffu, static_cast<unsigned>(found)));
      }
    }
  }
  return prefix;
}

// gap = min(gap, a - b) with unsigned wrap-around: one fused add-and-minimum (VIADDMNMX) for 32-bit keys
__device__ __forceinline__ void min_diff(unsigned int& gap, unsigned int a, unsigned int b) {
  gap = __viaddmin_u32(a, 0u - b, gap);
}
__device__ __forceinline__ void min_diff(unsigned long long& gap, unsigned long long a, unsigned long long b) {
  const unsigned long long d = a - b;
  gap = d < gap ? d : gap;
}

// c += (key < split): one compare and one predicated add, in place (the C++ form costs a register move per key)
__device__ __forceinline__ void count_below(int& c, unsigned int key, unsigned int split) {
  asm("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %1, %2;\n\t@p add.s32 %0, %0, 1;\n\t}" : "+r"(c) : "r"(key), "r"(split));
}
__device__ __forceinline__ void count_below(int& c, unsigned long long key, unsigned long long split) {
  asm("{\n\t.reg .pred p;\n\tsetp.lt.u64 p, %1, %2;\n\t@p add.s32 %0, %0, 1;\n\t}" : "+r"(c) : "l"(key), "l"(split));
}

// Same selection for keys held WITHOUT a validity mask: empty slots carry the all-ones sentinel (no finite, infinite
// or NaN input maps to it), `n` valid keys in the warp, 0 <= k < n.  The count of a round is taken against the
// absolute threshold `prefix | 1 << b` (one compare and one add per key); the candidates with the bit clear are that
// count minus the number of keys below the prefix, which earlier rounds already established.
template <typename K, int CAP>
__device__ __forceinline__ K warp_select_dense(const K (&keys)[CAP], int n, int k) {
  constexpr int BITS = sizeof(K) * 8;
  K prefix = 0;
  int below = 0;  // keys < prefix
  int cand = n;   // keys in [prefix, prefix + 2^(b+1))
  for (int b = BITS - 1; b >= 0; --b) {
    const K split = prefix | (static_cast<K>(1) << b);
    int c = 0, c1 = 0;  // two chains
#pragma unroll
    for (int i = 0; i < CAP; ++i) count_below((i & 1) ? c1 : c, keys[i], split);
    c = __reduce_add_sync(0xffffffffu, c + c1);
    if (k >= c) {
      cand -= c - below;
      below = c;
      prefix = split;
    } else {
      cand = c - below;
    }
    // Early exits: the answer is the SMALLEST key of the remaining range [prefix, prefix + 2^b) (rank k equals the
    // number of keys below the prefix) or its LARGEST key; both are found with one subtraction and one unsigned
    // minimum per key, no range test.  Smallest: keys below the prefix wrap around to differences >= 2^BITS - prefix,
    // larger than any genuine difference.  Largest: keys above the range wrap around to >= prefix + 2^b, keys below it
    // give >= 2^b, keys inside < 2^b; the top range also holds the sentinels and keeps bisecting instead.
    // This covers "one candidate left" and ends two to three rounds earlier on average.
    const int k_rel = k - below;
    if (b > 0 && (k_rel == 0 || k_rel == cand - 1)) {
      const K last = prefix | ((static_cast<K>(1) << b) - 1);
      const bool smallest = k_rel == 0;
      if (smallest || last != ~static_cast<K>(0)) {
        K gap = ~static_cast<K>(0);
        if (smallest) {
#pragma unroll
          for (int i = 0; i < CAP; ++i) {
            min_diff(gap, keys[i], prefix);
          }
        } else {
#pragma unroll
          for (int i = 0; i < CAP; ++i) {
            min_diff(gap, last, keys[i]);
          }
        }
        if constexpr (sizeof(K) == 8) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const K other = __shfl_xor_sync(0xffffffffu, gap, o);
            gap = other < gap ? other : gap;
          }
        } else {
          gap = static_cast<K>(__reduce_min_sync(0xffffffffu, static_cast<unsigned>(gap)));
        }
        return smallest ? prefix + gap : last - gap;
      }
    }
  }
  return prefix;
}

// ------------------------------------------------------------------------------ stats
template <typename T>
__device__ __forceinline__ void bitonic_sort_warp(T* a, int n_pow2, int lane) {
  for (int k = 2; k <= n_pow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < n_pow2; i += 32) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const T x0 = a[i], x1 = a[ixj];
          const bool up = (i & k) == 0;
          if ((x0 > x1) == up) { a[i] = x1; a[ixj] = x0; }
        }
      }
      __syncwarp();
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kRedThreads)
stat_kernel(const T* __restrict__ x, T* __restrict__ y, long long batch, long long xbs, long long xrs,
            long long ybs, Bins rows, Bins cols, int cols_per_chunk, int n_chunks, int op, int skipna,
            T fill, int tile_elems, int sort_cap) {
  extern __shared__ __align__(16) unsigned char smem[];
  T* tile = reinterpret_cast<T*>(smem);
  T* scratch = tile + (tile_elems + 3) / 4 * 4;  // [kRedWarps * sort_cap], median only
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long tasks = batch * rows.n_out * n_chunks;
  const int Wo = cols.n_out;
  const T nan = quiet_nan<T>();

  for (long long task = blockIdx.x; task < tasks; task += gridDim.x) {
    const int chunk = static_cast<int>(task % n_chunks);
    const int k = static_cast<int>((task / n_chunks) % rows.n_out);
    const long long b = task / (static_cast<long long>(n_chunks) * rows.n_out);
    const int l0 = chunk * cols_per_chunk;
    const int l1 = min(l0 + cols_per_chunk, Wo);
    T* yrow = y + b * ybs + static_cast<long long>(rows.out_index[k]) * Wo;
    const bool row_ok = rows.cover[k] != 0;
    const int r0 = rows.start[k], nr = row_ok ? rows.count[k] : 0;
    const int p0 = cols.start[l0];
    const int ncol = cols.start[l1 - 1] + cols.count[l1 - 1] - p0;
    const bool fits = static_cast<long long>(nr) * ncol <= tile_elems;  // host contract; never overrun
    if (fits && nr > 0 && ncol > 0) load_tile(tile, x + b * xbs, xrs, rows, cols, r0, nr, p0, ncol);
    __syncthreads();

    for (int l = l0 + warp; l < l1; l += kRedWarps) {
      T out = fill;
      if (fits && row_ok && cols.cover[l]) {
        const int c0 = cols.start[l] - p0, nc = cols.count[l];
        // pass 1: count, NaN count, sum, min, max
        double sum = 0.0;
        int n = 0, n_nan = 0;
        T vmin = CUDART_INF, vmax = -CUDART_INF;
        for (int r = 0; r < nr; ++r) {
          const T* trow = tile + r * ncol + c0;
          for (int c = lane; c < nc; c += 32) {
            const T v = trow[c];
            if (v == v) {
              sum += static_cast<double>(v);
              ++n;
              vmin = v < vmin ? v : vmin;
              vmax = v > vmax ? v : vmax;
            } else {
              ++n_nan;
            }
          }
        }
        sum = warp_sum(sum);
        n = warp_sum(n);
        n_nan = warp_sum(n_nan);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const T a = __shfl_xor_sync(0xffffffffu, vmin, o), c = __shfl_xor_sync(0xffffffffu, vmax, o);
          vmin = a < vmin ? a : vmin;
          vmax = c > vmax ? c : vmax;
        }
        const bool poisoned = !skipna && n_nan > 0;
        if (poisoned) {
          out = nan;
        } else if (n == 0) {
          out = (op == OP_SUM) ? T(0) : nan;
        } else if (op == OP_SUM) {
          out = static_cast<T>(sum);
        } else if (op == OP_MEAN) {
          out = static_cast<T>(sum / n);
        } else if (op == OP_MIN) {
          out = vmin;
        } else if (op == OP_MAX) {
          out = vmax;
        } else if (op == OP_VAR || op == OP_STD) {
          const double mean = sum / n;  // two passes, like np.var (ddof = 0)
          double m2 = 0.0;
          for (int r = 0; r < nr; ++r) {
            const T* trow = tile + r * ncol + c0;
            for (int c = lane; c < nc; c += 32) {
              const T v = trow[c];
              if (v == v) { const double d = static_cast<double>(v) - mean; m2 += d * d; }
            }
          }
          m2 = warp_sum(m2) / n;
          out = static_cast<T>(op == OP_STD ? sqrt(m2) : m2);
        } else if (nr <= kSelCap && nc <= 32) {
          // OP_MEDIAN, cells of up to 32 x 32: lane c keeps column c of the cell (one key per row) in
          // registers; bitwise radix select over the warp
          using K = decltype(to_key(T(0)));
          K keys[kSelCap];
          unsigned valid = 0;
#pragma unroll
          for (int q = 0; q < kSelCap; ++q) {
            T v = nan;
            if (q < nr && lane < nc) v = tile[q * ncol + c0 + lane];
            const bool ok = (v == v);
            keys[q] = ok ? to_key(v) : K(0);
            valid |= (ok ? 1u : 0u) << q;
          }
          const int k_lo = (n - 1) >> 1, k_hi = n >> 1;
          const K key_lo = warp_select<K, kSelCap>(keys, valid, k_lo);
          K key_hi = key_lo;
          if (k_hi != k_lo) {
            // the next order statistic: the same value if it occurs often enough, else the smallest larger one
            int le = 0;
            K next = ~K(0);
#pragma unroll
            for (int i = 0; i < kSelCap; ++i) {
              if ((valid >> i) & 1u) {
                le += keys[i] <= key_lo ? 1 : 0;
                if (keys[i] > key_lo && keys[i] < next) next = keys[i];
              }
            }
            le = __reduce_add_sync(0xffffffffu, le);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              const K other = __shfl_xor_sync(0xffffffffu, next, o);
              next = other < next ? other : next;
            }
            key_hi = (le > k_hi) ? key_lo : next;
          }
          const double lo = static_cast<double>(from_key(key_lo)), hi = static_cast<double>(from_key(key_hi));
          out = static_cast<T>((lo + hi) / 2);
        } else {  // OP_MEDIAN, larger cells: gather the valid values, sort, average the two middle ones
          T* s = scratch + warp * sort_cap;
          int n_pow2 = 1;
          while (n_pow2 < n) n_pow2 <<= 1;
          // stable compaction is not needed for a median: lanes claim slots with a warp prefix sum
          int base = 0;
          for (int r = 0; r < nr; ++r) {
            const T* trow = tile + r * ncol + c0;
            for (int cb = 0; cb < nc; cb += 32) {
              const int c = cb + lane;
              const T v = c < nc ? trow[c] : nan;
              const bool ok = (v == v);
              const unsigned m = __ballot_sync(0xffffffffu, ok);
              if (ok) s[base + __popc(m & ((1u << lane) - 1u))] = v;
              base += __popc(m);
            }
          }
          for (int i = n + lane; i < n_pow2; i += 32) s[i] = CUDART_INF;
          __syncwarp();
          bitonic_sort_warp(s, n_pow2, lane);
          const double lo = static_cast<double>(s[(n - 1) >> 1]), hi = static_cast<double>(s[n >> 1]);
          out = static_cast<T>((lo + hi) / 2);
          __syncwarp();
        }
      }
      if (lane == 0) yrow[cols.out_index[l]] = out;
    }
    __syncthreads();
  }
}

// ------------------------------------------------- TMA bulk staging (mbarrier completion)
__device__ __forceinline__ void red_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void red_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool red_mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void red_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  uint64_t pol;  // (source cells are read exactly once: evict-first in L2)
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "l"(pol)
      : "memory");
}

// ------------------------------------------------- mode, bit-sliced counters in registers
// For <= 32 dense labels the per-label counters of a lane's cell are kept BIT-SLICED in registers: plane b
// holds bit b of all 32 counters (bit j of the word = label j).  A row of the cell is turned into one-hot
// masks (1 << label), the masks are summed per bit position with a carry-save adder tree (LOP3 full adders:
// 3-input XOR and majority), and the resulting 6-bit column sums are added to the planes with one ripple
// pass.  No shared-memory read-modify-write, no atomics: ~3 ALU instructions per element.
__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xe8;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
__device__ __forceinline__ uint32_t onehot(uint32_t label) {
  uint32_t r;  // PTX shl clamps the shift amount: labels >= 32 give 0
  asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(1u), "r"(label));
  return r;
}

#include "csa_networks.inc"  // csa_sum<N>(m, s): generated Wallace trees (tools/gen_csa.py)

// Round 2: (1) byte labels get their one-hot mask from a 256-entry shared-memory table (an LDS on the
// otherwise idle load pipe instead of a shift on the ALU pipe, which bounds this kernel; the table also maps
// ANY list of <= 32 byte labels to positions, not just 0..n-1); (2) two rows of the cell feed ONE adder tree,
// so the ripple into the counter planes runs once per two rows; (3) PLANES = 10 when a cell holds <= 1023
// elements (12 otherwise).
constexpr int kPlanesMax = 12;  // counters up to 4095 per label and cell (the launcher checks the cell size)
constexpr int kModeTableBytes = 1024, kModeBarBytes = 128;

// one-hot masks of one row segment of this lane's cell (columns past the cell's width contribute nothing)
template <typename Tin, int NIN, bool RAGGED>
__device__ __forceinline__ void row_masks(const Tin* rowbase, const Tin* trow, int nc, const uint32_t* pad,
                                          const uint32_t* s_tab, uint32_t* m) {
  if constexpr (sizeof(Tin) == 1) {
    // byte labels: read the segment as aligned words, realign with funnel shifts, pick bytes at compile-time
    // positions and look the mask up; bytes outside the label list (and the 0xff padding) map to 0
    const int pos = static_cast<int>(trow - rowbase);
    const uint32_t* wrow = reinterpret_cast<const uint32_t*>(rowbase) + (pos >> 2);
    const int shb = (pos & 3) * 8;
    constexpr int NW = (NIN + 3) / 4;
    uint32_t wv[NW + 1];
#pragma unroll
    for (int q = 0; q <= NW; ++q) wv[q] = wrow[q];
    uint32_t al[NW];
#pragma unroll
    for (int q = 0; q < NW; ++q) {
      al[q] = __funnelshift_r(wv[q], wv[q + 1], shb);
      if constexpr (RAGGED) al[q] |= pad[q];  // cells narrower than NIN columns: the rest reads as 0xff
    }
#pragma unroll
    for (int c = 0; c < NIN; ++c) m[c] = s_tab[__byte_perm(al[c >> 2], 0, 0x4440 + (c & 3))];
  } else {
#pragma unroll
    for (int c = 0; c < NIN; ++c) {
      const Tin t = trow[c];
      const uint32_t v = (t >= Tin(0) && t < Tin(32)) ? static_cast<uint32_t>(t) : 32u;
      m[c] = (!RAGGED || c < nc) ? onehot(v) : 0u;
    }
  }
}

template <typename Tin, typename Tout, int NIN, int PLANES, bool RAGGED>
__global__ void __launch_bounds__(kLaneThreads, 3)
mode_planes_kernel(const Tin* __restrict__ x, Tout* __restrict__ y, long long batch, long long xbs, long long xrs,
                   long long ybs, Bins rows, Bins cols, const Tin* __restrict__ values, int n_values, int anti,
                   Tout fill, int stage_elems, int stage_rows) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // shared memory: [one-hot table 1 KB][2 mbarriers per warp][2 staging buffers per warp]
  // table: label byte -> one-hot of its position in `values` (0 for bytes that are not listed); it sits at offset
  // 0 so that a lookup is one LDS with a scaled index
  uint32_t* s_tab = reinterpret_cast<uint32_t*>(smem);
  const uint32_t bar0 = static_cast<uint32_t>(__cvta_generic_to_shared(smem + kModeTableBytes)) + 16u * warp;
  Tin* stage = reinterpret_cast<Tin*>(smem + kModeTableBytes + kModeBarBytes) +
               static_cast<size_t>(warp) * 2 * stage_rows * stage_elems;
  for (int i = tid; i < 256; i += kLaneThreads) s_tab[i] = 0u;
  if (lane == 0) {
    red_mbar_init(bar0, 1);
    red_mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if constexpr (sizeof(Tin) == 1) {
    if (tid == 0)
      for (int i = 0; i < n_values && i < 32; ++i) s_tab[static_cast<uint8_t>(values[i])] = 1u << i;
  }
  __syncthreads();
  uint32_t bar_uses = 0;  // bit b: parity of the next wait on this warp's buffer b
  const int Wo = cols.n_out;
  const int strips = (Wo + 31) / 32;
  const long long tasks = batch * rows.n_out * strips;
  const long long warp_id = static_cast<long long>(blockIdx.x) * kLaneWarps + warp;
  const long long n_warps = static_cast<long long>(gridDim.x) * kLaneWarps;
  const uint32_t label_mask = n_values >= 32 ? 0xffffffffu : ((1u << n_values) - 1u);

  for (long long task = warp_id; task < tasks; task += n_warps) {
    const int strip = static_cast<int>(task % strips);
    const int k = static_cast<int>((task / strips) % rows.n_out);
    const long long b = task / (static_cast<long long>(strips) * rows.n_out);
    const int l = strip * 32 + lane;
    const bool in = l < Wo;
    const int l_last = min(strip * 32 + 31, Wo - 1);
    const bool row_ok = rows.cover[k] != 0;
    const int r0 = rows.start[k], nr = row_ok ? rows.count[k] : 0;
    const int p0 = cols.start[strip * 32];
    const int ncol = cols.start[l_last] + cols.count[l_last] - p0;  // host contract: <= stage_elems
    const int c0 = in ? cols.start[l] - p0 : 0;
    const int nc = (in && cols.cover[l]) ? cols.count[l] : 0;       // host contract: <= NIN
    const Tin* xb = x + b * xbs;
    const int s0 = src_index(cols, p0);
    const bool contiguous = cols.src_step == 1 && s0 + ncol <= cols.n_src;
    const int row_first = src_index(rows, r0);

    uint32_t plane[PLANES];
#pragma unroll
    for (int q = 0; q < PLANES; ++q) plane[q] = 0;
    // byte labels: columns past this lane's cell width are forced to 0xff (one-hot of 255 = 0), one OR per
    // word instead of a compare + select per element
    constexpr int kPadWords = (NIN + 3) / 4;
    uint32_t pad[kPadWords];
#pragma unroll
    for (int q = 0; q < kPadWords; ++q) {
      uint32_t w = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) w |= (4 * q + e >= nc) ? (0xffu << (8 * e)) : 0u;
      pad[q] = w;
    }

    // double-buffered staging: chunk ch + 1 is in flight (LDGSTS) while chunk ch is counted
    const int n_chunks = (nr + stage_rows - 1) / stage_rows;
    // Contiguous strips (the usual case): rows are issued and consumed in order, so the segment pointer and its
    // 16-byte phase are carried along (pointer += row stride, phase = (phase + stride) mod 16) instead of being
    // rebuilt from the row index with 64-bit multiplies for every row (that cost 60 of 230 instructions per row).
    const long long rstride = xrs * rows.src_step;  // elements between two consecutive rows of the cell
    const int rsb16 = static_cast<int>((rstride * static_cast<long long>(sizeof(Tin))) & 15);
    const Tin* seg_next = xb + static_cast<long long>(row_first) * xrs + s0;  // segment of the next row to issue
    int pb_next = static_cast<int>(reinterpret_cast<uintptr_t>(seg_next) & 15);  // its phase in bytes
    int pb_cons = pb_next;                                                        // phase of the next row consumed
    const int seg_bytes = ncol * static_cast<int>(sizeof(Tin));
    const uint32_t stage_s = static_cast<uint32_t>(__cvta_generic_to_shared(stage));
    auto issue = [&](int ch) {
      Tin* buf = stage + (ch & 1) * stage_rows * stage_elems;
      const int rb = ch * stage_rows, nrb = min(stage_rows, nr - rb);
      if (contiguous) {
        // TMA: lane r copies row r of the chunk with ONE bulk copy (the 16-byte aligned range enclosing its
        // segment: source - phase .. segment end rounded up), all rows of the chunk in one instruction; the
        // copies complete on the buffer's mbarrier.  (The per-16-byte cp.async loop this replaces cost 57
        // instructions per row, a fifth of the kernel.)
        const uint32_t bar = bar0 + 8u * (ch & 1);
        const int pb = (pb_next + lane * rsb16) & 15;
        const uint32_t bytes = lane < nrb ? static_cast<uint32_t>((pb + seg_bytes + 15) & ~15) : 0u;
        const uint32_t total = __reduce_add_sync(0xffffffffu, bytes);
        if (lane == 0) red_mbar_expect_tx(bar, total);
        __syncwarp();
        if (lane < nrb) {
          const char* a = reinterpret_cast<const char*>(seg_next + static_cast<long long>(lane) * rstride) - pb;
          red_bulk_load(stage_s + static_cast<uint32_t>(((ch & 1) * stage_rows + lane) * stage_elems * sizeof(Tin)),
                        a, bytes, bar);
        }
        seg_next += static_cast<long long>(nrb) * rstride;
        pb_next = (pb_next + nrb * rsb16) & 15;
        return;
      }
      for (int r = 0; r < nrb; ++r) {
        const Tin* src = xb + static_cast<long long>(row_first + rows.src_step * (rb + r)) * xrs;
        stage_row_async<Tin>(buf + r * stage_elems, src, cols, p0, s0, ncol, contiguous, lane);
      }
    };
    if (n_chunks > 0) issue(0);
    cp_async_commit();
    for (int ch = 0; ch < n_chunks; ++ch) {
      if (ch + 1 < n_chunks) issue(ch + 1);
      cp_async_commit();
      if (contiguous) {
        const uint32_t b = ch & 1;
        while (!red_mbar_try_wait(bar0 + 8u * b, (bar_uses >> b) & 1u)) {}
        bar_uses ^= 1u << b;
      } else {
        cp_async_wait<1>();
      }
      __syncwarp();
      const Tin* buf = stage + (ch & 1) * stage_rows * stage_elems;
      const int rb = ch * stage_rows, nrb = min(stage_rows, nr - rb);
      for (int r = 0; r < nrb; r += 2) {
        const bool two = r + 1 < nrb;
        int sh0 = 0, sh1 = 0;  // the rows' 16-byte phases (see stage_row)
        if (contiguous) {
          sh0 = pb_cons / static_cast<int>(sizeof(Tin));
          pb_cons = (pb_cons + rsb16) & 15;
          if (two) {
            sh1 = pb_cons / static_cast<int>(sizeof(Tin));
            pb_cons = (pb_cons + rsb16) & 15;
          }
        }
        const Tin* base0 = buf + r * stage_elems;
        uint32_t sum[7];
        if (two) {  // two rows of the cell through one carry-save adder tree: 7-bit sums per label position
          uint32_t m[2 * NIN];
          row_masks<Tin, NIN, RAGGED>(base0, base0 + sh0 + c0, nc, pad, s_tab, m);
          row_masks<Tin, NIN, RAGGED>(base0 + stage_elems, base0 + stage_elems + sh1 + c0, nc, pad, s_tab, m + NIN);
          csa_sum<2 * NIN>(m, sum);
        } else {
          uint32_t m[NIN];
          row_masks<Tin, NIN, RAGGED>(base0, base0 + sh0 + c0, nc, pad, s_tab, m);
          csa_sum<NIN>(m, sum);
        }
        // ripple the column sums into the planes
        uint32_t carry = 0;
#pragma unroll
        for (int q = 0; q < PLANES; ++q) {
          const uint32_t add = q < 7 ? sum[q] : 0u;
          const uint32_t nc_ = maj3(plane[q], add, carry);
          plane[q] = xor3(plane[q], add, carry);
          carry = nc_;
        }
      }
      __syncwarp();  // the buffer is refilled by the next iteration's issue
    }
    // ---- decode without transposing: walk the planes from the top bit down, keeping the labels that still
    // tie for the largest (smallest) count; the lowest surviving bit is the first label (idxmax / idxmin
    // tie rule).  A zero count is the reference's -1: still the smallest value, and "all zero" keeps every
    // label, so label 0 wins exactly like idxmax over an all -1 column.
    Tout out = fill;
    if (row_ok && in && cols.cover[l]) {
      uint32_t cand = label_mask;
#pragma unroll
      for (int q = PLANES - 1; q >= 0; --q) {
        const uint32_t t = cand & (anti ? ~plane[q] : plane[q]);
        cand = t ? t : cand;
      }
      out = static_cast<Tout>(values[__ffs(cand) - 1]);
    }
    if (in) y[b * ybs + static_cast<long long>(rows.out_index[k]) * Wo + cols.out_index[l]] = out;
  }
}

// ------------------------------------------------------------ stats, one cell per lane
// sum / mean / min / max / var / std with the strip staging of the mode kernel: a warp owns 32 adjacent
// target cells, every lane accumulates its own cell in registers (fp64 sums), no shuffles, no atomics.
// With 25-column cells the lanes' shared-memory reads are conflict-free (stride 25 words, gcd(25, 32) = 1).
// Rows are staged with LDGSTS into two buffers per warp (the next chunk is in flight while this one is
// reduced); columns are walked in unrolled blocks of 8, branch-free (NaN -> predicated add, FMNMX skips NaN).
// var / std of fp64 data make a second pass over the strip around the exact mean, like np.var; fp32 data
// use one pass with shifted fp64 moments (see below).
// OPC selects the accumulator set at compile time: 0 = sum / mean, 1 = min / max, 2 = var / std.
template <typename T, int OPC>
__global__ void __launch_bounds__(kLaneThreads, 3)
stat_lanes_kernel(const T* __restrict__ x, T* __restrict__ y, long long batch, long long xbs, long long xrs,
                  long long ybs, Bins rows, Bins cols, int op, int skipna, T fill, int stage_elems,
                  int stage_rows) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  T* stage = reinterpret_cast<T*>(smem) + static_cast<size_t>(warp) * 2 * stage_rows * stage_elems;
  const int Wo = cols.n_out;
  const int strips = (Wo + 31) / 32;
  const long long tasks = batch * rows.n_out * strips;
  const long long warp_id = static_cast<long long>(blockIdx.x) * kLaneWarps + warp;
  const long long n_warps = static_cast<long long>(gridDim.x) * kLaneWarps;
  const T nan = quiet_nan<T>();

  for (long long task = warp_id; task < tasks; task += n_warps) {
    const int strip = static_cast<int>(task % strips);
    const int k = static_cast<int>((task / strips) % rows.n_out);
    const long long b = task / (static_cast<long long>(strips) * rows.n_out);
    const int l = strip * 32 + lane;
    const bool in = l < Wo;
    const int l_last = min(strip * 32 + 31, Wo - 1);
    const bool row_ok = rows.cover[k] != 0;
    const int r0 = rows.start[k], nr = row_ok ? rows.count[k] : 0;
    const int p0 = cols.start[strip * 32];
    const int ncol = cols.start[l_last] + cols.count[l_last] - p0;  // host contract: <= stage_elems
    const int c0 = in ? cols.start[l] - p0 : 0;
    const int nc = (in && cols.cover[l]) ? cols.count[l] : 0;
    const T* xb = x + b * xbs;
    const int s0 = src_index(cols, p0);
    const bool contiguous = cols.src_step == 1 && s0 + ncol <= cols.n_src;
    const int row_first = src_index(rows, r0);
    const int n_chunks = (nr + stage_rows - 1) / stage_rows;

    auto issue = [&](int ch) {
      T* buf = stage + (ch & 1) * stage_rows * stage_elems;
      const int rb = ch * stage_rows, nrb = min(stage_rows, nr - rb);
      for (int r = 0; r < nrb; ++r) {
        const T* src = xb + static_cast<long long>(row_first + rows.src_step * (rb + r)) * xrs;
        stage_row_async<T>(buf + r * stage_elems, src, cols, p0, s0, ncol, contiguous, lane);
      }
    };
    // walk every staged row segment of this lane's cell: row(trow, nc) once per segment, f(v) per element
    // 16-byte phase of a contiguous row segment, carried from row to row (rows are consumed in order)
    const int rsb16 = static_cast<int>((xrs * rows.src_step * static_cast<long long>(sizeof(T))) & 15);
    const int pb0 = static_cast<int>(
        reinterpret_cast<uintptr_t>(xb + static_cast<long long>(row_first) * xrs + s0) & 15);
    auto sweep = [&](auto&& row, auto&& f) {
      int pb = pb0;
      if (n_chunks > 0) issue(0);
      cp_async_commit();
      for (int ch = 0; ch < n_chunks; ++ch) {
        if (ch + 1 < n_chunks) issue(ch + 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncwarp();
        const T* buf = stage + (ch & 1) * stage_rows * stage_elems;
        const int rb = ch * stage_rows, nrb = min(stage_rows, nr - rb);
        for (int r = 0; r < nrb; ++r) {
          int sh = 0;  // the row's 16-byte phase (see stage_row)
          if (contiguous) {
            sh = pb / static_cast<int>(sizeof(T));
            pb = (pb + rsb16) & 15;
          }
          const T* trow = buf + r * stage_elems + sh + c0;
          row(trow, nc);
          int c = 0;
          for (; c + 8 <= nc; c += 8) {
            T v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = trow[c + e];
#pragma unroll
            for (int e = 0; e < 8; ++e) f(v[e]);
          }
          for (; c < nc; ++c) f(trow[c]);
        }
        __syncwarp();  // the buffer is refilled by the next iteration's issue
      }
    };

    double sum = 0.0, m2 = 0.0, mean = 0.0;
    int n = 0;
    T vmin = CUDART_INF, vmax = -CUDART_INF;
    auto no_row = [](const T*, int) {};
    if constexpr (OPC == 1) {
      sweep(no_row, [&](T v) {
        n += (v == v) ? 1 : 0;
        vmin = v < vmin ? v : vmin;  // comparisons with NaN are false: NaN never replaces a value
        vmax = v > vmax ? v : vmax;
      });
    } else if constexpr (OPC == 2 && sizeof(T) == 4) {
      // fp32 data: ONE pass with fp64 moments around the cell's first valid value.  The shift is a sample
      // of the cell, so (sum d)^2 / n <= 2 n var and the cancellation costs at most ~2n * 2^-53 relative:
      // far below the fp32 rounding of the result (the strip is ~300 MB in flight across the GPU, a
      // second pass would come from DRAM again).
      bool have = false;
      double shift = 0.0;
      float shift_f = 0.f;
      auto find_shift = [&](const T* trow, int ncols) {  // once per row segment until a sample is found
        if (have) return;
        for (int c = 0; c < ncols; ++c) {
          const T v = trow[c];
          if (v == v) { shift_f = static_cast<float>(v); shift = static_cast<double>(v); have = true; break; }
        }
      };
      sweep(find_shift, [&](T v) {
        // a NaN is replaced by the shift itself BEFORE the conversion (d = 0 exactly): one fp32 select instead of a
        // 64-bit one, and the count is one predicated add on the same predicate (the C++ form costs a move more)
        float vs;
        asm("{\n\t.reg .pred p;\n\tsetp.num.f32 p, %2, %2;\n\t@p add.s32 %0, %0, 1;\n\tselp.f32 %1, %2, %3, p;\n\t}"
            : "+r"(n), "=f"(vs)
            : "f"(static_cast<float>(v)), "f"(shift_f));
        const double d = static_cast<double>(vs) - shift;
        sum += d;
        m2 += d * d;
      });
      if (n > 0) {
        const double md = sum / n;
        m2 = fmax(m2 - md * sum, 0.0);
        mean = shift + md;
      }
    } else {
      sweep(no_row, [&](T v) {
        const bool ok = v == v;
        n += ok ? 1 : 0;
        sum += ok ? static_cast<double>(v) : 0.0;
      });
      if (n > 0) mean = sum / n;
      if constexpr (OPC == 2) {
        sweep(no_row, [&](T v) {
          const double d = (v == v) ? static_cast<double>(v) - mean : 0.0;
          m2 += d * d;
        });
      }
    }
    const int n_nan = nr * nc - n;
    T out = fill;
    if (row_ok && in && cols.cover[l]) {
      const bool poisoned = !skipna && n_nan > 0;
      if (poisoned) out = nan;
      else if (n == 0) out = (op == OP_SUM) ? T(0) : nan;
      else if (op == OP_SUM) out = static_cast<T>(sum);
      else if (op == OP_MEAN) out = static_cast<T>(mean);
      else if (op == OP_MIN) out = vmin;
      else if (op == OP_MAX) out = vmax;
      else if (op == OP_VAR) out = static_cast<T>(m2 / n);
      else out = static_cast<T>(sqrt(m2 / n));
    }
    if (in) y[b * ybs + static_cast<long long>(rows.out_index[k]) * Wo + cols.out_index[l]] = out;
  }
}

// ------------------------------------------------------------ median, one warp per cell, no staging
// Cells of up to CAP rows x 32 columns: lane c loads column c of the cell straight from global memory (one key
// per row, all rows in flight, 4 or 8 bytes per lane: a row segment is one coalesced request and neighbouring
// cells share cache lines through L1 / L2), converts to order-preserving keys and runs the bitwise select
// in registers.  No shared memory, no CTA barrier: warps are independent and many are resident.
// CTAs of 4 warps, 7 per SM: 72 registers per thread allow 28 resident warps (3 x 8 warps would stop at 24)
constexpr int kMedThreads = 128;
__device__ __forceinline__ int src_index_near(const Bins& a, int pos) {
  const int i = a.src_offset + a.src_step * pos;
  return static_cast<unsigned>(i) < static_cast<unsigned>(a.n_src) ? i : src_index(a, pos);  // modulo only on a wrap
}

template <typename T, int CAP>
__global__ void __launch_bounds__(kMedThreads, sizeof(T) == 8 ? 5 : 8)  // 64-bit keys: 96 registers, 20 warps
median_direct_kernel(const T* __restrict__ x, T* __restrict__ y, long long batch, long long xbs, long long xrs,
                     long long ybs, Bins rows, Bins cols, int skipna, T fill) {
  const int lane = threadIdx.x & 31;
  // 32-bit cell arithmetic: the launcher keeps batch * Ho * Wo + the grid's warps below 2^32 per launch
  const unsigned warp_id = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned n_warps = (gridDim.x * blockDim.x) >> 5;
  const int Ho = rows.n_out, Wo = cols.n_out;
  const unsigned cells = static_cast<unsigned>(batch) * Ho * Wo;
  const T nan = quiet_nan<T>();
  using K = decltype(to_key(T(0)));
  // (b, k, l) of the warp's cell advance by a fixed step with carries: no division per cell
  const unsigned step_rows = n_warps / Wo;
  const int dl = static_cast<int>(n_warps - step_rows * Wo);
  const unsigned db = step_rows / Ho;
  const int dk = static_cast<int>(step_rows - db * Ho);
  int l = static_cast<int>(warp_id % Wo);
  int k = static_cast<int>((warp_id / Wo) % Ho);
  unsigned b = warp_id / Wo / Ho;
  for (unsigned cell = warp_id; cell < cells; cell += n_warps, l += dl, k += dk, b += db) {
    if (l >= Wo) { l -= Wo; ++k; }
    if (k >= Ho) { k -= Ho; ++b; }
    T out = fill;
    if (rows.cover[k] != 0 && cols.cover[l] != 0) {
      const int nr = rows.count[k], nc = cols.count[l];  // host contract: nr <= CAP, nc <= 32
      const T* col = x + b * xbs + src_index_near(cols, cols.start[l] + (lane < nc ? lane : 0));
      const int r_first = src_index_near(rows, rows.start[k]);
      const int r_last = r_first + rows.src_step * (nr - 1);
      const int nr_own = lane < nc ? nr : 0;  // lanes outside the cell load nothing
      K keys[CAP];  // empty slots (outside the cell, NaN) hold the all-ones sentinel
      int n_own = 0;
      if (r_last >= 0 && r_last < rows.n_src) {
        // the usual case, the cell's source rows do not wrap around the axis: one pointer increment per row
        const T* ptr = col + static_cast<long long>(r_first) * xrs;
        const long long stride = xrs * rows.src_step;
#pragma unroll
        for (int q = 0; q < CAP; ++q) {
          T v = nan;
          if (q < nr_own) v = __ldg(ptr);
          ptr += stride;
          const bool ok = (v == v);
          keys[q] = ok ? to_key(v) : ~K(0);
          n_own += ok ? 1 : 0;
        }
      } else {
        int ri = r_first;
#pragma unroll
        for (int q = 0; q < CAP; ++q) {
          T v = nan;
          if (q < nr_own) v = __ldg(col + static_cast<long long>(ri) * xrs);
          ri += rows.src_step;  // next source row, modulo n_src
          ri = ri >= rows.n_src ? ri - rows.n_src : (ri < 0 ? ri + rows.n_src : ri);
          const bool ok = (v == v);
          keys[q] = ok ? to_key(v) : ~K(0);
          n_own += ok ? 1 : 0;
        }
      }
      const int n = __reduce_add_sync(0xffffffffu, n_own);
      if ((!skipna && n < nr * nc) || n == 0) {
        out = nan;
      } else {
        const int k_lo = (n - 1) >> 1, k_hi = n >> 1;
        const K key_lo = warp_select_dense<K, CAP>(keys, n, k_lo);
        K key_hi = key_lo;
        if (k_hi != k_lo) {
          // the next order statistic: the same value if it occurs often enough, else the smallest larger one
          // (the sentinel is never <= a valid key).  Smallest larger key: keys <= key_lo wrap around to
          // differences larger than any genuine one, so one subtraction and one minimum per key find it.
          int le = 0;
          const K above = key_lo + 1;  // key_lo is a valid key, hence not the all-ones sentinel
          K gap = ~K(0);
#pragma unroll
          for (int i = 0; i < CAP; ++i) {
            count_below(le, keys[i], above);
            min_diff(gap, keys[i], above);
          }
          K next = above + gap;
          le = __reduce_add_sync(0xffffffffu, le);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const K other = __shfl_xor_sync(0xffffffffu, next, o);
            next = other < next ? other : next;
          }
          key_hi = (le > k_hi) ? key_lo : next;
        }
        const double lo = static_cast<double>(from_key(key_lo)), hi = static_cast<double>(from_key(key_hi));
        out = static_cast<T>((lo + hi) / 2);
      }
    }
    if (lane == 0) y[b * ybs + static_cast<long long>(rows.out_index[k]) * Wo + cols.out_index[l]] = out;
  }
}

// --------------------------------------------------------------------------- launchers
struct ChunkPlan {
  int cols_per_chunk, n_chunks, tile_elems;
};

template <typename Tin, typename Tout>
int mode_launch(const void* x, void* y, int64_t batch, int64_t xbs, int64_t xrs, int64_t ybs,
                const rg_bins_t* rows_c, const rg_bins_t* cols_c, int32_t cols_per_chunk, int32_t tile_elems,
                int32_t strip_elems, int32_t max_bin_rows, int32_t max_bin_cols, const void* values_sorted,
                const int32_t* value_rank, const void* values, int32_t n_values, int32_t dense, int32_t anti,
                double fill, cudaStream_t stream) {
  Bins rows, cols;
  int rc = make_bins(rows_c, &rows);
  if (rc != RG_OK) return rc;
  rc = make_bins(cols_c, &cols);
  if (rc != RG_OK) return rc;
  DeviceInfo dev;
  rc = device_info(&dev);
  if (rc != RG_OK) return rc;
  // ---- fastest path: <= 32 labels (dense 0..n-1, or ANY list when the labels are bytes: the kernel's lookup
  //      table maps them), cells of <= 32 columns: bit-sliced counters in registers
  const long long cell_pop = static_cast<long long>(max_bin_rows) * max_bin_cols;
  if ((dense || sizeof(Tin) == 1) && n_values <= 32 && strip_elems > 0 && max_bin_cols > 0 && max_bin_cols <= 32 &&
      max_bin_rows > 0 && cell_pop <= (1 << kPlanesMax) - 1) {
    const int stage_elems = (strip_elems + 48 + 15) / 16 * 16;
    // two staging buffers per warp; three CTAs per SM share the shared memory; an even number of rows per buffer
    // (the kernel feeds row pairs into one adder tree)
    int stage_rows = static_cast<int>((56 * 1024) / (static_cast<size_t>(kLaneWarps) * 2 * stage_elems * sizeof(Tin)));
    if (stage_rows > 8) stage_rows = 8;
    if (stage_rows > 1) stage_rows &= ~1;
    if (stage_rows >= 1) {
      const size_t smem_p = static_cast<size_t>(kLaneWarps) * 2 * stage_rows * stage_elems * sizeof(Tin) +
                            kModeTableBytes + kModeBarBytes;
      using PK = void (*)(const Tin*, Tout*, long long, long long, long long, long long, Bins, Bins, const Tin*,
                          int, int, Tout, int, int);
      const bool few = cell_pop <= 1023;
      // every covered cell exactly N columns wide (regular grids whose width has an instance): no per-word
      // padding masks in the kernel
#define RG_MP(N)                                                                                             \
  (few ? (cols_c->uniform_count != N ? mode_planes_kernel<Tin, Tout, N, 10, true>                            \
                                     : mode_planes_kernel<Tin, Tout, N, 10, false>)                          \
       : (cols_c->uniform_count != N ? mode_planes_kernel<Tin, Tout, N, 12, true>                            \
                                     : mode_planes_kernel<Tin, Tout, N, 12, false>))
      PK kp = max_bin_cols <= 8    ? RG_MP(8)
              : max_bin_cols <= 16 ? RG_MP(16)
              : max_bin_cols <= 25 ? RG_MP(25)
              : max_bin_cols <= 26 ? RG_MP(26)
                                   : RG_MP(32);
#undef RG_MP
      if (smem_p > 48 * 1024)
        RG_CUDA(cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_p)));
      int per_sm_p = 0;
      RG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_p, kp, kLaneThreads, smem_p));
      if (per_sm_p >= 1) {
        const long long strips = (cols.n_out + 31) / 32;
        const long long tasks_p = static_cast<long long>(batch) * rows.n_out * strips;
        long long grid_p = static_cast<long long>(dev.sm_count) * per_sm_p * 2;
        const long long need = (tasks_p + kLaneWarps - 1) / kLaneWarps;
        if (grid_p > need) grid_p = need;
        kp<<<static_cast<unsigned>(grid_p), kLaneThreads, smem_p, stream>>>(
            static_cast<const Tin*>(x), static_cast<Tout*>(y), batch, xbs, xrs, ybs, rows, cols,
            static_cast<const Tin*>(values), n_values, anti, static_cast<Tout>(fill), stage_elems, stage_rows);
        return static_cast<int>(cudaGetLastError());
      }
    }
  }
  // ---- throughput path: one cell per lane, lane-private counters (needs few labels and strips whose
  //      source columns fit the per-warp staging buffer)
  if (n_values <= kLaneMaxLabels && strip_elems > 0) {
    const int stage_elems = (strip_elems + 48 + 15) / 16 * 16;  // slack: 16-byte phase + whole-word reads
    const size_t smem_l = static_cast<size_t>(n_values + 1) * kLaneThreads * 4 +
                          static_cast<size_t>(kLaneWarps) * kStageRows * stage_elems * sizeof(Tin);
    if (smem_l <= 100 * 1024) {
      auto kl = mode_lanes_kernel<Tin, Tout>;
      if (smem_l > 48 * 1024)
        RG_CUDA(cudaFuncSetAttribute(kl, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_l)));
      int per_sm_l = 0;
      RG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_l, kl, kLaneThreads, smem_l));
      if (per_sm_l >= 1) {
        const long long strips = (cols.n_out + 31) / 32;
        const long long tasks_l = static_cast<long long>(batch) * rows.n_out * strips;
        long long grid_l = static_cast<long long>(dev.sm_count) * per_sm_l * 2;
        const long long need = (tasks_l + kLaneWarps - 1) / kLaneWarps;
        if (grid_l > need) grid_l = need;
        kl<<<static_cast<unsigned>(grid_l), kLaneThreads, smem_l, stream>>>(
            static_cast<const Tin*>(x), static_cast<Tout*>(y), batch, xbs, xrs, ybs, rows, cols,
            static_cast<const Tin*>(values_sorted), value_rank, static_cast<const Tin*>(values), n_values,
            dense, anti, static_cast<Tout>(fill), stage_elems);
        return static_cast<int>(cudaGetLastError());
      }
    }
  }
  const size_t smem = (static_cast<size_t>(tile_elems) * sizeof(Tin) + 15) / 16 * 16 +
                      static_cast<size_t>(kRedWarps) * n_values * sizeof(int32_t);
  if (smem > static_cast<size_t>(dev.max_smem_optin)) return RG_ERR_SMEM;
  auto kern = mode_kernel<Tin, Tout>;
  if (smem > 48 * 1024)
    RG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const int n_chunks = (cols.n_out + cols_per_chunk - 1) / cols_per_chunk;
  const long long tasks = static_cast<long long>(batch) * rows.n_out * n_chunks;
  int per_sm = 0;
  RG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kRedThreads, smem));
  if (per_sm < 1) return RG_ERR_SMEM;
  long long grid = static_cast<long long>(dev.sm_count) * per_sm * 4;
  if (grid > tasks) grid = tasks;
  kern<<<static_cast<unsigned>(grid), kRedThreads, smem, stream>>>(
      static_cast<const Tin*>(x), static_cast<Tout*>(y), batch, xbs, xrs, ybs, rows, cols, cols_per_chunk,
      n_chunks, static_cast<const Tin*>(values_sorted), value_rank, static_cast<const Tin*>(values), n_values,
      dense, anti, static_cast<Tout>(fill), tile_elems);
  return static_cast<int>(cudaGetLastError());
}

template <typename T>
int stat_launch(const T* x, T* y, int64_t batch, int64_t xbs, int64_t xrs, int64_t ybs,
                const rg_bins_t* rows_c, const rg_bins_t* cols_c, int32_t cols_per_chunk, int32_t tile_elems,
                int32_t op, int32_t skipna, int32_t sort_cap, int32_t strip_elems, int32_t max_bin_rows,
                int32_t max_bin_cols, double fill, cudaStream_t stream) {
  if (!x || !y) return RG_ERR_NULL;
  if (op < OP_SUM || op > OP_MEDIAN || cols_per_chunk <= 0 || tile_elems <= 0) return RG_ERR_SHAPE;
  Bins rows, cols;
  int rc = make_bins(rows_c, &rows);
  if (rc != RG_OK) return rc;
  rc = make_bins(cols_c, &cols);
  if (rc != RG_OK) return rc;
  if (batch <= 0) return batch == 0 ? RG_OK : RG_ERR_SHAPE;
  DeviceInfo dev;
  rc = device_info(&dev);
  if (rc != RG_OK) return rc;
  // ---- median of cells up to 32 x 32: one warp per cell, keys in registers, no staging
  if (op == OP_MEDIAN && max_bin_rows > 0 && max_bin_rows <= 32 && max_bin_cols > 0 && max_bin_cols <= 32) {
    using MK = void (*)(const T*, T*, long long, long long, long long, long long, Bins, Bins, int, T);
    MK mk = max_bin_rows <= 8    ? median_direct_kernel<T, 8>
            : max_bin_rows <= 16 ? median_direct_kernel<T, 16>
            : max_bin_rows <= 25 ? median_direct_kernel<T, 25>
                                 : median_direct_kernel<T, 32>;
    int per_sm_m = 0;
    RG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_m, mk, kMedThreads, 0));
    if (per_sm_m >= 1) {
      // the kernel counts cells in 32 bits: slabs are dealt to launches of < 2^31 cells each
      const long long slab_cells = static_cast<long long>(rows.n_out) * cols.n_out;
      if (slab_cells < (1ll << 31)) {
        const long long b_max = std::max<long long>(1, ((1ll << 31) - 1) / slab_cells);
        for (long long b0 = 0; b0 < batch; b0 += b_max) {
          const long long nb = std::min<long long>(b_max, batch - b0);
          const long long cells = nb * slab_cells;
          long long grid_m = static_cast<long long>(dev.sm_count) * per_sm_m * 4;
          const long long need = (cells + kMedThreads / 32 - 1) / (kMedThreads / 32);
          if (grid_m > need) grid_m = need;
          mk<<<static_cast<unsigned>(grid_m), kMedThreads, 0, stream>>>(x + b0 * xbs, y + b0 * ybs, nb, xbs, xrs, ybs,
                                                                        rows, cols, skipna, static_cast<T>(fill));
          const cudaError_t err = cudaGetLastError();
          if (err != cudaSuccess) return static_cast<int>(err);
        }
        return RG_OK;
      }
    }
  }
  // ---- throughput path (everything but the median): one cell per lane
  if (op != OP_MEDIAN && strip_elems > 0) {
    const int stage_elems = (strip_elems + 48 + 15) / 16 * 16;
    // two staging buffers per warp; aim at three CTAs per SM, but a single row per buffer is enough
    int stage_rows = static_cast<int>((72 * 1024) / (static_cast<size_t>(kLaneWarps) * 2 * stage_elems * sizeof(T)));
    if (stage_rows > 8) stage_rows = 8;
    if (stage_rows < 1) stage_rows = 1;
    if (static_cast<size_t>(kLaneWarps) * 2 * stage_rows * stage_elems * sizeof(T) <=
        static_cast<size_t>(dev.max_smem_optin)) {
      const size_t smem_l = static_cast<size_t>(kLaneWarps) * 2 * stage_rows * stage_elems * sizeof(T);
      auto kl = (op == OP_MIN || op == OP_MAX)   ? stat_lanes_kernel<T, 1>
                : (op == OP_VAR || op == OP_STD) ? stat_lanes_kernel<T, 2>
                                                 : stat_lanes_kernel<T, 0>;
      if (smem_l > 48 * 1024)
        RG_CUDA(cudaFuncSetAttribute(kl, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_l)));
      int per_sm_l = 0;
      RG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_l, kl, kLaneThreads, smem_l));
      if (per_sm_l >= 1) {
        const long long strips = (cols.n_out + 31) / 32;
        const long long tasks_l = static_cast<long long>(batch) * rows.n_out * strips;
        long long grid_l = static_cast<long long>(dev.sm_count) * per_sm_l * 2;
        const long long need = (tasks_l + kLaneWarps - 1) / kLaneWarps;
        if (grid_l > need) grid_l = need;
        kl<<<static_cast<unsigned>(grid_l), kLaneThreads, smem_l, stream>>>(
            x, y, batch, xbs, xrs, ybs, rows, cols, op, skipna, static_cast<T>(fill), stage_elems, stage_rows);
        return static_cast<int>(cudaGetLastError());
      }
    }
  }
  const size_t smem = (static_cast<size_t>((tile_elems + 3) / 4 * 4) +
                       (op == OP_MEDIAN ? static_cast<size_t>(kRedWarps) * sort_cap : 0)) * sizeof(T);
  if (smem > static_cast<size_t>(dev.max_smem_optin)) return RG_ERR_SMEM;
  auto kern = stat_kernel<T>;
  if (smem > 48 * 1024)
    RG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const int n_chunks = (cols.n_out + cols_per_chunk - 1) / cols_per_chunk;
  const long long tasks = static_cast<long long>(batch) * rows.n_out * n_chunks;
  int per_sm = 0;
  RG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kRedThreads, smem));
  if (per_sm < 1) return RG_ERR_SMEM;
  long long grid = static_cast<long long>(dev.sm_count) * per_sm * 4;
  if (grid > tasks) grid = tasks;
  kern<<<static_cast<unsigned>(grid), kRedThreads, smem, stream>>>(x, y, batch, xbs, xrs, ybs, rows, cols,
                                                                  cols_per_chunk, n_chunks, op, skipna,
                                                                  static_cast<T>(fill), tile_elems, sort_cap);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace rg

extern "C" {

int rg_mode(const void* x, void* y, int32_t in_type, int32_t out_type, int64_t batch,
            int64_t x_batch_stride, int64_t x_row_stride, int64_t y_batch_stride, const rg_bins_t* rows,
            const rg_bins_t* cols, int32_t cols_per_chunk, int32_t tile_elems, int32_t strip_elems,
            int32_t max_bin_rows, int32_t max_bin_cols, const void* values_sorted, const int32_t* value_rank,
            const void* values, int32_t n_values, int32_t dense, int32_t anti_mode, double fill, void* stream) {
  if (!x || !y || !values || (!dense && (!values_sorted || !value_rank))) return RG_ERR_NULL;
  if (n_values <= 0 || cols_per_chunk <= 0 || tile_elems <= 0 || batch < 0) return RG_ERR_SHAPE;
  if (batch == 0) return RG_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define RG_MODE(CODE, TIN)                                                                                  \
  if (in_type == CODE) {                                                                                    \
    if (out_type == CODE)                                                                                   \
      return rg::mode_launch<TIN, TIN>(x, y, batch, x_batch_stride, x_row_stride, y_batch_stride, rows,     \
                                       cols, cols_per_chunk, tile_elems, strip_elems, max_bin_rows,         \
                                       max_bin_cols,                                                        \
                                       values_sorted, value_rank, values, n_values, dense, anti_mode, fill, \
                                       s);                                                                  \
    if (out_type == RG_F64)                                                                                 \
      return rg::mode_launch<TIN, double>(x, y, batch, x_batch_stride, x_row_stride, y_batch_stride, rows,  \
                                          cols, cols_per_chunk, tile_elems, strip_elems, max_bin_rows,      \
                                          max_bin_cols,                                                     \
                                          values_sorted, value_rank, values, n_values, dense, anti_mode,    \
                                          fill, s);                                                         \
    return RG_ERR_UNSUPPORTED;                                                                              \
  }
  RG_MODE(RG_U8, uint8_t)
  RG_MODE(RG_I8, int8_t)
  RG_MODE(RG_I16, int16_t)
  RG_MODE(RG_I32, int32_t)
  RG_MODE(RG_I64, long long)
#undef RG_MODE
  return RG_ERR_UNSUPPORTED;
}

int rg_stat_f64(const double* x, double* y, int64_t batch, int64_t x_batch_stride, int64_t x_row_stride,
                int64_t y_batch_stride, const rg_bins_t* rows, const rg_bins_t* cols,
                int32_t cols_per_chunk, int32_t tile_elems, int32_t op, int32_t skipna, int32_t sort_cap,
                int32_t strip_elems, int32_t max_bin_rows, int32_t max_bin_cols, double fill, void* stream) {
  return rg::stat_launch<double>(x, y, batch, x_batch_stride, x_row_stride, y_batch_stride, rows, cols,
                                 cols_per_chunk, tile_elems, op, skipna, sort_cap, strip_elems, max_bin_rows,
                                 max_bin_cols, fill, static_cast<cudaStream_t>(stream));
}

int rg_stat_f32(const float* x, float* y, int64_t batch, int64_t x_batch_stride, int64_t x_row_stride,
                int64_t y_batch_stride, const rg_bins_t* rows, const rg_bins_t* cols,
                int32_t cols_per_chunk, int32_t tile_elems, int32_t op, int32_t skipna, int32_t sort_cap,
                int32_t strip_elems, int32_t max_bin_rows, int32_t max_bin_cols, double fill, void* stream) {
  return rg::stat_launch<float>(x, y, batch, x_batch_stride, x_row_stride, y_batch_stride, rows, cols,
                                cols_per_chunk, tile_elems, op, skipna, sort_cap, strip_elems, max_bin_rows,
                                max_bin_cols, fill, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
