/*
 * regrid_b200.h -- C ABI of the B200-native regridding engine.
 *
 * The reference (xarray-regrid v0.4.2) is pure Python with no FFI; the narrowest
 * seams a native engine plugs into are the three array-math calls its methods
 * delegate to (reference paths relative to src/xarray_regrid/):
 *
 *   apply_weights(da, weights, skipna, nan_threshold)   methods/conservative.py:181-208
 *   data.interp(coords=..., method=...)                  methods/interp.py:43-46
 *   flox.xarray.xarray_reduce(...)                       methods/flox_reduce.py:89-96, 174-182
 *
 * Every entry point below replaces one of those calls (cited per function).
 * Conventions:
 *   - plain pointers and sizes only; no C++/torch types cross this boundary;
 *   - "device" pointers are CUDA device memory of the current device, owned by the
 *     caller; the library never allocates or frees caller-visible memory;
 *   - calls are asynchronous and ordered on `stream` (a cudaStream_t passed as void*;
 *     NULL = the legacy default stream); they are re-entrant and keep no global state;
 *   - return value: 0 = RG_OK, negative = argument error (RG_ERR_*), positive = the
 *     cudaError_t of the failing runtime call.  No exceptions cross the boundary.
 *   - data layout: [batch, rows(lat), cols(lon)], unit stride along cols; batch and row
 *     strides are given in ELEMENTS.  All non-spatial dims are flattened into `batch`.
 */
#ifndef REGRID_B200_H
#define REGRID_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RG_OK 0
#define RG_ERR_NULL (-1)      /* required pointer is NULL                         */
#define RG_ERR_SHAPE (-2)     /* non-positive / inconsistent extent or stride      */
#define RG_ERR_SMEM (-3)      /* problem does not fit the kernel's shared memory   */
#define RG_ERR_UNSUPPORTED (-4)
#define RG_ERR_ALIGN (-5)

#define RG_ABI_VERSION 4 /* 3: rg_pipeline_*, rg_host_copy, pole source rows in rg_row_nanmean_* / rg_conservative_host_*,
                            rg_conservative_*(skipna = 0) counts NaN as 0 like the reference's fillna(0)
                            4: rg_row_ring_t block layout (lane_cols == 2: strip_b0, block_l0, halo_l, halo_r) */

/* ABI version of the loaded library (== RG_ABI_VERSION of the header it was built from). */
int rg_abi_version(void);
/* Static, human-readable text for a return code (RG_ERR_* or a cudaError_t). */
const char* rg_error_string(int code);
/* Number of SMs of the current device (grid sizing is done inside the library). */
int rg_device_sm_count(void);

/*
 * Banded per-axis weight table ("ELL", tap-major).
 *
 * Replaces the dense [n_src, n_tgt] weight matrix of utils.py:145-169 /
 * methods/conservative.py:219-244 (get_weights, apply_spherical_correction,
 * format_weights): only structurally non-zero entries are kept, like the
 * reference's sparse.COO weights (conservative.py:297-300).
 *
 *   idx[t * n_out + o]   source index of tap t of output o (0 <= idx < n_src), or
 *                        n_src + {0,1} for the synthetic south / north pole row of
 *                        utils.py:313-336 (rows axis only, needs `pole` below)
 *   w  [t * n_out + o]   its weight, float type = the data's (float for f32, double for f64)
 *   cnt[o]               number of taps of output o (taps t >= cnt[o] are ignored)
 *   cover[o]             1 if the target centre lies inside the source centres' range
 *                        (conservative.py:123-125), 0 -> the output is NaN
 */
typedef struct rg_band {
  const int32_t* idx;   /* device, [k_max * n_out] */
  const void* w;        /* device, [k_max * n_out] */
  const int32_t* cnt;   /* device, [n_out]         */
  const uint8_t* cover; /* device, [n_out]         */
  int32_t n_out;
  int32_t k_max;
  int32_t n_src;
  int32_t span; /* max over outputs of (largest - smallest tap index + 1); 0 = unknown */
  int32_t window_rows; /* rows axis of an interpolation band: the largest multiple of 4 (<= 32) such that
                          every aligned group of that many consecutive outputs takes its taps from at most
                          16 consecutive source indices; 0 = no such grouping / unknown.  Lets the library
                          pick the shared-memory upsampling kernel (it stages 16 source rows per group). */
  int32_t group_span;  /* columns axis of an interpolation band whose taps are contiguous (idx[t] = idx[0] + t):
                          max over aligned groups of 4 consecutive covered outputs of (largest - smallest tap
                          index + 1); 0 = unknown / taps not contiguous.  <= 6 lets the 4-tap upsampling kernel
                          load one window per group instead of 4 taps per output. */
} rg_band_t;

/*
 * Optional row-ring schedule for the TMA-pipelined conservative kernel.
 *
 * The fast kernel is persistent and warp-specialised: a producer thread streams whole
 * source rows into a shared-memory ring with cp.async.bulk (TMA) + mbarriers and the
 * consumer warps contract them, so every source row is read from HBM once per run even
 * though neighbouring target rows share rows.  The schedule tells the producer in which
 * order to load rows and the consumers where each tap lives.  It is derived from the rows
 * band on the host (tables.row_ring_schedule); outputs are grouped into `n_runs`
 * contiguous runs, one (slab, run) pair = one unit of work for a CTA.
 *
 *   sched_row[p]              source row loaded at schedule position p
 *   run_first[r], run_sched[r]   first output / first schedule position of run r
 *                             (arrays of n_runs + 1 entries)
 *   krec[4 * o + 0..3]        per output row o, one 16-byte record:
 *                               [0] need     positions (relative to the run's first) that must
 *                                            have landed before o is computed
 *                               [1] release  relative positions that are dead once o is computed
 *                               [2] kcnt     taps of o (<= 8), or -1 when o is not covered (NaN row)
 *                               [3] tap_back 8 x 4 bits: where tap t lives, in rows back from the
 *                                            newest landed row (= need - 1 - its relative position)
 *   wrec[8 * o + t]           weight of tap t of output row o (element type of the data, zero
 *                             padded): the rows band's weights, output-major
 *   max_live                  max over o of rows simultaneously alive (ring must hold them)
 *
 * Column decomposition: the target columns are cut into `n_strips` strips of <= 32 (panel layouts: 64) outputs;
 * consumer warp w of `n_warps` handles strips w, w + n_warps, ...  For strip s
 *   strip_l0[s] .. strip_l0[s+1]-1   its target columns
 *   strip_c0[s], strip_nc[s]         the source columns (a range modulo n_src that may wrap
 *                                    around the date line) covering every tap of those targets;
 *                                    both are multiples of 16 / sizeof(element)
 *   strip_cols_max                   max over s of strip_nc[s]
 *
 * Column panels (two-phase kernel): the strips are grouped into `n_panels` panels of consecutive strips; a CTA
 * works on ONE panel of a (slab, run) unit, its ring rows hold only the panel's source columns
 *   panel_c0[p] .. panel_c0[p] + panel_nc[p] - 1   (modulo n_src; two bulk copies when the range wraps)
 * and consumer warp w owns strip panel_strip0[p] + w (n_warps = the largest number of strips in a panel).
 * strip_c0 is RELATIVE to the panel's first column.  Short panel rows let several CTAs -- independent row
 * pipelines -- share an SM, which is what grids with few source rows per target row (coarsening ratios of
 * 2 - 2.5) and rows wider than one CTA's shared memory (0.1 degree) need.  One panel covering the whole row
 * (panel_c0 = 0, panel_nc = n_src) is the classic layout; the register-resident kernel requires it.
 *
 * Passing NULL selects the generic kernel (any band, pole rows, unaligned rows).
 */
typedef struct rg_row_ring {
  const int32_t* sched_row; /* device, [n_sched]          */
  const int32_t* krec;      /* device, [4 * n_out], 16-byte aligned */
  const void* wrec;         /* device, [8 * n_out], 16-byte aligned */
  const int32_t* run_first; /* device, [n_runs + 1]       */
  const int32_t* run_sched; /* device, [n_runs + 1]       */
  const int32_t* strip_l0;  /* device, [n_strips + 1]     */
  const int32_t* strip_c0;  /* device, [n_strips]         */
  const int32_t* strip_nc;  /* device, [n_strips]         */
  int32_t n_runs;
  int32_t n_sched;
  int32_t max_live;
  int32_t pad_shift; /* shared-memory column padding: slot(j) = j + (j >> pad_shift); 31 = none */
  int32_t n_strips;
  int32_t strip_cols_max;
  int32_t n_warps;   /* consumer warps (1..24); the kernel adds one producer warp */
  int32_t lane_cols; /* 4: every target column takes taps idx0 .. idx0 + cnt - 1 (cnt <= 5, mod n_src), idx0 is
                        even and grows by exactly 4 from one target to the next inside a strip, every strip has
                        <= 31 targets, all covered, and n_strips <= n_warps: the kernel then keeps both
                        contractions in registers (one lane per target column).
                        2: block layout (coarsening by 2 - 3, rg_conservative_* only): a lane owns the 4 source
                        columns strip_c0[s] + 4 j .. + 3 (strip_c0 ABSOLUTE and even; a multiple of 4 for fp32
                        when n_panels > 1) of block strip_b0[s] + j and produces the targets block_l0[b] ..
                        block_l0[b + 1] - 1 (at most two) whose contiguous taps all lie within halo_l columns
                        before / halo_r columns after the block; <= 30 blocks per strip, all targets covered.
                        Panels: panel_c0[p] = first own column of the panel - 4, panel_nc[p] = 4 * (blocks + 2),
                        or one panel {0, n_src}; strips per panel <= n_warps <= 12.
                        0: no such structure. */
  const int32_t* panel_strip0; /* device, [n_panels + 1] */
  const int32_t* panel_c0;     /* device, [n_panels]     */
  const int32_t* panel_nc;     /* device, [n_panels]     */
  int32_t n_panels;            /* >= 1 */
  int32_t panel_cols_max;      /* max over p of panel_nc[p] */
  int32_t strip_targets_max;   /* max over s of strip_l0[s+1] - strip_l0[s]: <= 32, or <= 64 in panel layouts
                                  of 6 consumer warps (the kernel then walks a strip's targets in two passes) */
  const int32_t* strip_b0;     /* device, [n_strips + 1] or NULL: first block of every strip (lane_cols == 2) */
  const int32_t* block_l0;     /* device, [n_blocks + 1] or NULL: first target of every block (lane_cols == 2) */
  int32_t halo_l;              /* 1..2 (lane_cols == 2) */
  int32_t halo_r;              /* 1..2 (lane_cols == 2) */
} rg_row_ring_t;

/*
 * Fused conservative regrid of a [batch, H, W] stack onto [batch, Ho, Wo]:
 *
 *   valid[b,k,l] = sum_ij notnull(x[b,i,j]) * Wlat[i,k] * Wlon[j,l]
 *   out  [b,k,l] = sum_ij fillna0(x[b,i,j]) * Wlat[i,k] * Wlon[j,l]
 *   skipna:  y = out / valid where valid >= valid_thr else NaN
 *   !skipna: y = out -- a NaN counts as 0: the reference contracts da.fillna(0) whether or not
 *            skipna is set (conservative.py:196-198); only the division and the threshold are
 *            conditional (conservative.py:200-202).  (NaN-propagating contraction: rg_separable_*.)
 *   y = NaN where !(cover_lat[k] && cover_lon[l])
 *
 * Replaces apply_weights (methods/conservative.py:181-208) plus the coverage mask of
 * conservative.py:161-164 in ONE pass over x (no fillna/notnull/divide/where passes).
 * `valid_thr` is get_valid_threshold(nan_threshold) (conservative.py:211-216), computed
 * by the caller.  `pole` is NULL or a device [batch, 2, W] array whose rows hold the
 * longitude-mean of the first / last source row repeated W times (rg_row_nanmean_*),
 * referenced by idx == H / H+1: the synthetic pole rows are ordinary, 16-byte aligned rows.
 * H = rows.n_src, W = cols.n_src, Ho = rows.n_out, Wo = cols.n_out.  y is dense
 * [batch, Ho, Wo] with batch stride y_batch_stride.  `ring` is NULL or the row-ring
 * schedule above; the library silently uses the generic kernel when the ring kernel's
 * preconditions (16-byte aligned rows, <= 8 taps per axis, ring fits shared memory) do not hold.
 */
int rg_conservative_f64(const double* x, double* y, int64_t batch,
                        int64_t x_batch_stride, int64_t x_row_stride, int64_t y_batch_stride,
                        const rg_band_t* rows, const rg_band_t* cols,
                        int32_t skipna, double valid_thr, const double* pole,
                        const rg_row_ring_t* ring, void* stream);
int rg_conservative_f32(const float* x, float* y, int64_t batch,
                        int64_t x_batch_stride, int64_t x_row_stride, int64_t y_batch_stride,
                        const rg_band_t* rows, const rg_band_t* cols,
                        int32_t skipna, float valid_thr, const float* pole,
                        const rg_row_ring_t* ring, void* stream);

/*
 * Longitude-mean (NaN-skipping) of the southernmost and northernmost row of every slab, written as full rows:
 * pole[b,0,:] = nanmean(x[b,south_row,:]), pole[b,1,:] = nanmean(x[b,north_row,:]); all-NaN -> NaN.
 * pole is [batch, 2, W].  south_row / north_row are positions in the caller's (possibly descending or
 * unsorted) latitude order: the reference sorts first (ensure_monotonic, utils.py:277) and then takes
 * isel(lat=0) / isel(lat=-1).  Replaces the `.mean(lon_coord)` of format_lat (utils.py:320-333).
 */
int rg_row_nanmean_f64(const double* x, double* pole, int64_t batch, int32_t H, int32_t W,
                       int64_t x_batch_stride, int64_t x_row_stride, int32_t south_row, int32_t north_row,
                       void* stream);
int rg_row_nanmean_f32(const float* x, float* pole, int64_t batch, int32_t H, int32_t W,
                       int64_t x_batch_stride, int64_t x_row_stride, int32_t south_row, int32_t north_row,
                       void* stream);

/*
 * Persistent host <-> device streaming pipeline: three non-blocking streams (H2D, compute, D2H) and one
 * event triple per slot, created once and reused for any number of chunks and calls.  It stands in for the
 * reference's dask chunk scheduling (methods/conservative.py:263-302: weights chunked 1:1 with the data;
 * interp.py:43 and flox_reduce.py:89 run per dask block).  A "slot" is a caller-owned set of device buffers
 * (input chunk, result chunk, scratch); the pipeline only orders the work on them:
 *
 *   rg_pipeline_h2d(p, slot, ...)       the H2D stream waits until the kernels that read the slot's previous
 *                                       input are done, then ONE cudaMemcpyAsync
 *   rg_pipeline_compute_begin(p, slot)  the compute stream waits for the slot's H2D and for the D2H of the
 *                                       slot's previous result; the caller then launches any rg_* calls on
 *                                       rg_pipeline_stream(p, 1)
 *   rg_pipeline_compute_end(p, slot)    marks the slot's kernels as queued
 *   rg_pipeline_d2h(p, slot, ...)       the D2H stream waits for the slot's kernels, then ONE cudaMemcpyAsync
 *   rg_pipeline_sync(p, stage, slot)    host-blocks until the slot's last H2D (stage 0), kernels (1) or D2H (2)
 *                                       has finished (e.g. before a pageable staging buffer is refilled)
 *   rg_pipeline_wait_stream(p, stream)  all pipeline streams wait for the work queued so far on the caller's
 *                                       stream (buffers produced / recycled there are then safe to use)
 *   rg_pipeline_drain(p)                host-blocks until all three streams are idle
 *
 * Host pointers of rg_pipeline_h2d / _d2h should be page-locked for full PCIe bandwidth.  For pageable arrays
 * (numpy) rg_pipeline_h2d_pageable / _d2h_pageable cut the chunk into pieces that pass through a caller-owned
 * page-locked staging buffer (`stage_pinned`, split into 4 pieces; >= 16 KiB) with a multi-threaded memcpy
 * (rg_host_copy, `n_threads` threads) overlapping the DMA of the neighbouring piece -- still one
 * cudaMemcpyAsync per piece.  _h2d_pageable blocks the host only when it laps the staging ring;
 * _d2h_pageable returns when the whole chunk is in `dst_host` (queue the next chunk's work before it).
 * Not thread-safe per handle.
 */
typedef struct rg_pipeline rg_pipeline_t;
int rg_pipeline_create(rg_pipeline_t** out, int32_t n_slots); /* 1 <= n_slots <= 8 */
int rg_pipeline_destroy(rg_pipeline_t* p);                   /* drains first; NULL is a no-op */
int32_t rg_pipeline_slots(const rg_pipeline_t* p);
void* rg_pipeline_stream(rg_pipeline_t* p, int32_t which);   /* 0: H2D, 1: compute, 2: D2H (cudaStream_t) */
int rg_pipeline_h2d(rg_pipeline_t* p, int32_t slot, void* dst_device, const void* src_host, size_t bytes);
int rg_pipeline_compute_begin(rg_pipeline_t* p, int32_t slot);
int rg_pipeline_compute_end(rg_pipeline_t* p, int32_t slot);
int rg_pipeline_d2h(rg_pipeline_t* p, int32_t slot, void* dst_host, const void* src_device, size_t bytes);
int rg_pipeline_h2d_pageable(rg_pipeline_t* p, int32_t slot, void* dst_device, const void* src_host, size_t bytes,
                             void* stage_pinned, size_t stage_bytes, int32_t n_threads);
int rg_pipeline_d2h_pageable(rg_pipeline_t* p, int32_t slot, void* dst_host, const void* src_device, size_t bytes,
                             void* stage_pinned, size_t stage_bytes, int32_t n_threads);
int rg_pipeline_sync(rg_pipeline_t* p, int32_t stage, int32_t slot);
int rg_pipeline_wait_stream(rg_pipeline_t* p, void* stream);
int rg_pipeline_drain(rg_pipeline_t* p);
int rg_host_copy(void* dst, const void* src, size_t bytes, int32_t n_threads);

/*
 * Host-buffer variant of rg_conservative_*: x_host / y_host are HOST pointers
 * (pinned memory gives full PCIe bandwidth; pageable memory works but is staged by the
 * driver).  The stack is cut into chunks of `chunk_batch` slabs that are copied in,
 * regridded and copied out on the three streams of `pipeline` so that H2D, the kernel and D2H
 * overlap (stands in for the reference's dask chunk scheduling,
 * methods/conservative.py:263-302).  `pipeline` is a persistent rg_pipeline_t, or NULL: a
 * temporary 3-slot pipeline is created and destroyed inside the call.  `workspace` is
 * caller-owned DEVICE memory of at least rg_conservative_host_workspace(..., n_slots) bytes
 * (n_slots = rg_pipeline_slots(pipeline), 3 for NULL).  The call returns after the last D2H
 * copy has completed, unless RG_HOST_NO_DRAIN is set in `flags` (persistent pipelines only): the
 * caller then queues further calls and finishes with rg_pipeline_drain.  After an error the
 * pipeline is drained before returning: nothing stays in flight into the caller's buffers.
 * pole_rows != 0: the rows band references the synthetic pole rows; south_row / north_row as in
 * rg_row_nanmean_*.  Host layout is dense: x_host [batch, H, W], y_host [batch, Ho, Wo].
 */
#define RG_HOST_NO_DRAIN 1
size_t rg_conservative_host_workspace(int64_t chunk_batch, int32_t H, int32_t W, int32_t Ho,
                                      int32_t Wo, int32_t elem_size, int32_t n_slots);
int rg_conservative_host_f64(const double* x_host, double* y_host, int64_t batch,
                             int64_t chunk_batch, const rg_band_t* rows, const rg_band_t* cols,
                             int32_t skipna, double valid_thr, int32_t pole_rows, int32_t south_row,
                             int32_t north_row, const rg_row_ring_t* ring, void* workspace,
                             size_t workspace_bytes, rg_pipeline_t* pipeline, int32_t flags);
int rg_conservative_host_f32(const float* x_host, float* y_host, int64_t batch,
                             int64_t chunk_batch, const rg_band_t* rows, const rg_band_t* cols,
                             int32_t skipna, float valid_thr, int32_t pole_rows, int32_t south_row,
                             int32_t north_row, const rg_row_ring_t* ring, void* workspace,
                             size_t workspace_bytes, rg_pipeline_t* pipeline, int32_t flags);

/*
 * Separable banded contraction y[b,k,l] = sum_t sum_c wr[t,k] * wc[c,l] * x[b, ri[t,k], ci[c,l]] with NaN
 * propagating through every tap (weight 0 included) and y = NaN where !(cover_r[k] && cover_c[l]).
 * Same kernels as rg_conservative_*, without any NaN bookkeeping.  It replaces `data.interp(coords, method)`
 * (methods/interp.py:43-46 -> scipy interp1d per axis) for
 *   linear: 2-tap bands (tables.linear_band);
 *   cubic : first the prefilter bands (A^-1 of the not-a-knot collocation system, source grid ->
 *           formatted source grid), then the 4-tap B-spline evaluation bands (tables.cubic_bands);
 *   nearest on floating-point data: 1-tap bands (integers: rg_nearest below).
 */
int rg_separable_f64(const double* x, double* y, int64_t batch,
                     int64_t x_batch_stride, int64_t x_row_stride, int64_t y_batch_stride,
                     const rg_band_t* rows, const rg_band_t* cols, const double* pole,
                     const rg_row_ring_t* ring, void* stream);
int rg_separable_f32(const float* x, float* y, int64_t batch,
                     int64_t x_batch_stride, int64_t x_row_stride, int64_t y_batch_stride,
                     const rg_band_t* rows, const rg_band_t* cols, const float* pole,
                     const rg_row_ring_t* ring, void* stream);

/* Element type codes for the type-generic entry points. */
#define RG_U8 0
#define RG_I8 1
#define RG_I16 2
#define RG_I32 3
#define RG_I64 4
#define RG_F32 5
#define RG_F64 6

/*
 * Nearest-neighbour regrid as a pure index gather (bit-exact):
 *   y[b,k,l] = (out_type) x[b, row_idx[k], col_idx[l]]   where row_ok[k] && col_ok[l], else `fill`.
 * row_idx / col_idx are computed on the host with scipy's own expression
 * (searchsorted(x[1:]/2 + x[:-1]/2, t, side="left"): ties go to the lower neighbour), *_ok marks
 * targets inside [x[0], x[-1]].  out_type is in_type, or RG_F64 (the reference returns float64 with NaN
 * for integer input: scipy/interpolate/_interpolate.py:336-338).  Replaces data.interp(method="nearest")
 * (methods/interp.py:43-46).
 */
int rg_nearest(const void* x, void* y, int32_t in_type, int32_t out_type, int64_t batch,
               int64_t x_batch_stride, int64_t x_row_stride, int64_t y_batch_stride,
               int32_t Ho, int32_t Wo, const int32_t* row_idx, const uint8_t* row_ok,
               const int32_t* col_idx, const uint8_t* col_ok, double fill, void* stream);

/*
 * Spline coefficients of the cubic method, in place: solves A c = y along `axis` (0: rows, 1: columns) of
 * every slab of the dense-column [batch, H, W] fp64 stack, where A = L U is the banded (two sub-, two
 * super-diagonals) not-a-knot collocation matrix of scipy's make_interp_spline(x, y, k=3)
 * (scipy/interpolate/_bsplines.py; the reference reaches it through xarray's interp ->
 * scipy.interpolate.interp1d(kind="cubic"), methods/interp.py:43-46).  The caller factors A without
 * pivoting and passes lower = [l1 | l2] ([2, n]: L[i, i-1], L[i, i-2]) and upper = [1/d | u1 | u2]
 * ([3, n]: 1/U[i, i], U[i, i+1], U[i, i+2]), n = H or W, device arrays.  RG_ERR_UNSUPPORTED when the
 * factors do not fit shared memory (n > ~5800): use the truncated-inverse band pass then.
 */
int rg_spline_solve_f64(double* y, int64_t batch, int64_t y_batch_stride, int64_t y_row_stride, int32_t H,
                        int32_t W, int32_t axis, const double* lower, const double* upper, void* stream);

/*
 * The cubic spline's NaN rule (scipy interp1d: one NaN anywhere in the array makes the WHOLE result
 * NaN, _interpolate.py:409-437): rg_nan_flag_* ORs 1 into *flag (device int32, zeroed by the caller)
 * if any element of the [batch, H, W] stack is NaN; rg_poison_if_* overwrites y[0..n) with NaN when
 * *flag != 0.  Both are stream-ordered; the flag never visits the host.
 */
int rg_nan_flag_f64(const double* x, int64_t batch, int32_t H, int32_t W, int64_t x_batch_stride,
                    int64_t x_row_stride, int32_t* flag, void* stream);
int rg_nan_flag_f32(const float* x, int64_t batch, int32_t H, int32_t W, int64_t x_batch_stride,
                    int64_t x_row_stride, int32_t* flag, void* stream);
int rg_poison_if_f64(double* y, int64_t n, const int32_t* flag, void* stream);
int rg_poison_if_f32(float* y, int64_t n, const int32_t* flag, void* stream);

/*
 * Per-axis bin table of the group-by regridders.  Bin membership is decided on the host with the
 * reference's expressions (methods/_shared.py:12-18: closed-left bins on
 * `append(t, t[-1] + step) - step / 2`; _shared.py:92-108 trimming), on the source axis sorted ascending
 * and, for longitude, shifted / wrap-padded by format_for_regrid(stats=True) (utils.py:258-261, 341-390).
 * A bin is therefore a contiguous range of "formatted positions":
 *   start[o], count[o]   the positions of target o (bins are visited in ascending order of o)
 *   cover[o]             target centre inside the (trimmed) source range (_shared.py:56-59), else fill
 *   out_index[o]         index of target o in the output (the target's own order: reindex_like)
 *   source index of position p = (src_offset + src_step * p) mod n_src,  src_step = +1 or -1
 */
typedef struct rg_bins {
  const int32_t* start;     /* device, [n_out] */
  const int32_t* count;     /* device, [n_out] */
  const uint8_t* cover;     /* device, [n_out] */
  const int32_t* out_index; /* device, [n_out] */
  int32_t n_out;
  int32_t n_src;
  int32_t src_offset;
  int32_t src_step;
  int32_t uniform_count; /* the count of EVERY covered bin when they are all equal (regular grids), else 0:
                            lets the register-resident mode kernel drop its per-cell width masks */
} rg_bins_t;

/*
 * most_common / least_common: for every target cell count the occurrences of each label of `values`
 * among the source cells of its bin and return the label with the largest (anti_mode: smallest) count.
 * Labels that do not occur count as -1 (the reference passes fill_value=-1 to flox), ties go to the
 * label that comes first in `values`, labels not listed in `values` are ignored; cells that are not
 * covered get `fill`.  Replaces the flox count + idxmax/idxmin of methods/flox_reduce.py:174-183 and
 * the masking of methods/_shared.py:41-73.
 *   values         device [n_values], element type in_type, in the caller's order
 *   dense != 0     values == 0..n_values-1: a label is its own index (values_sorted / value_rank unused)
 *   values_sorted  device [n_values] ascending, value_rank[i] = position of values_sorted[i] in `values`
 *   cols_per_chunk target columns staged per CTA; tile_elems >= rows-per-bin * source columns of any chunk
 *   strip_elems    max source columns spanned by 32 consecutive target columns (0 = unknown): enables the
 *                  one-cell-per-lane throughput kernel when n_values <= 64
 *   max_bin_rows, max_bin_cols   largest number of source rows / columns in one target cell (0 = unknown):
 *                  with dense labels, n_values <= 32, cells of <= 32 columns and <= 4095 source cells the
 *                  counters live bit-sliced in registers (carry-save adder kernel)
 *   in_type        RG_U8 / RG_I8 / RG_I16 / RG_I32 / RG_I64; out_type = in_type or RG_F64 (NaN fill)
 */
int rg_mode(const void* x, void* y, int32_t in_type, int32_t out_type, int64_t batch,
            int64_t x_batch_stride, int64_t x_row_stride, int64_t y_batch_stride,
            const rg_bins_t* rows, const rg_bins_t* cols, int32_t cols_per_chunk, int32_t tile_elems,
            int32_t strip_elems, int32_t max_bin_rows, int32_t max_bin_cols, const void* values_sorted,
            const int32_t* value_rank, const void* values, int32_t n_values, int32_t dense, int32_t anti_mode,
            double fill, void* stream);

/* Statistic codes of rg_stat_*. */
#define RG_STAT_SUM 0
#define RG_STAT_MEAN 1
#define RG_STAT_VAR 2
#define RG_STAT_STD 3
#define RG_STAT_MIN 4
#define RG_STAT_MAX 5
#define RG_STAT_MEDIAN 6

/*
 * Zonal statistic per target cell over the source cells of its bin: sum, mean, var / std (population,
 * ddof = 0, two passes), min, max, median.  skipna == 0: a NaN in the bin makes the result NaN;
 * skipna != 0: NaNs are ignored (an all-NaN bin gives NaN, 0 for sum).  Cells that are not covered get
 * `fill`.  Replaces flox.xarray.xarray_reduce(data, *coords, func=method, expected_groups=bounds,
 * skipna=..., fill_value=...) of methods/flox_reduce.py:89-96.  sort_cap (median only) >= next power
 * of two of the largest bin population; strip_elems as in rg_mode (0 = unknown: tile kernel only);
 * max_bin_rows / max_bin_cols as in rg_mode (0 = unknown): cells of <= 32 x 32 take the one-warp-per-cell
 * register median kernel.
 */
int rg_stat_f64(const double* x, double* y, int64_t batch, int64_t x_batch_stride, int64_t x_row_stride,
                int64_t y_batch_stride, const rg_bins_t* rows, const rg_bins_t* cols,
                int32_t cols_per_chunk, int32_t tile_elems, int32_t op, int32_t skipna, int32_t sort_cap,
                int32_t strip_elems, int32_t max_bin_rows, int32_t max_bin_cols, double fill, void* stream);
int rg_stat_f32(const float* x, float* y, int64_t batch, int64_t x_batch_stride, int64_t x_row_stride,
                int64_t y_batch_stride, const rg_bins_t* rows, const rg_bins_t* cols,
                int32_t cols_per_chunk, int32_t tile_elems, int32_t op, int32_t skipna, int32_t sort_cap,
                int32_t strip_elems, int32_t max_bin_rows, int32_t max_bin_cols, double fill, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* REGRID_B200_H */
