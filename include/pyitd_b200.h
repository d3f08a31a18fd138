/*
 * pyitd_b200.h -- C ABI of the B200-native ITD sifting loop (libpyitd_b200.so).
 *
 * The reference (falseywinchnet/PyITD) has no FFI layer: its boundary for this path is one Python
 * class and two numba functions in /root/reference/ITD.py.  Each entry point below names the
 * reference interface it replaces; INTEGRATION.md shows the ctypes binding a maintainer of the
 * reference would add to route ITD.itd through this library.
 *
 * Conventions
 *   - plain pointers and sizes only; `stream` is a cudaStream_t passed as void* (NULL = default
 *     stream); every `_device` call is asynchronous on that stream, every `_host` call returns
 *     after the result is in the caller's host buffers;
 *   - signals are rows of a C-contiguous (n_signals, n_samples) matrix;
 *   - outputs are C-contiguous (n_signals, rows, n_samples) with rows = max_iteration + 2
 *     (the reference hard-codes 22 rows, ITD.py:384-385; ITD_numba.py:102-103 intends
 *     max_iteration + 1 rotations + 1 trend row);
 *   - every function returns 0 on success or a negative PYITD_E_* code; per-signal algorithmic
 *     conditions (the exceptions the reference raises) are reported in `status[signal]`.
 */
#ifndef PYITD_B200_H
#define PYITD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PYITD_ABI_VERSION 1

#if defined(__GNUC__)
#define PYITD_API __attribute__((visibility("default")))
#else
#define PYITD_API
#endif

/* precision variants */
#define PYITD_F64        0   /* float64 in/out, float64 carry + arithmetic: the reference's dtype   */
#define PYITD_F32_MIXED  1   /* float32 in/out, float64 carry + arithmetic == float32(reference)    */
#define PYITD_F32        2   /* float32 in/out/carry/arithmetic (pure fp32)          */

/* per-signal status bits (status[signal]) */
#define PYITD_ST_OK        0
#define PYITD_ST_ZERO_DX   1  /* X[k+1] == X[k] in a segment: reference raises ZeroDivisionError, ITD.py:116 */
#define PYITD_ST_NONFINITE 2  /* NaN/Inf in the input: the reference's NaN path (ITD.py:46-51,64-68) is unsupported */
#define PYITD_ST_TOO_SHORT 4  /* fewer than 3 samples: undefined in the reference, ITD.py:42-43 */

/* stop kinds (stop_kind[signal]) */
#define PYITD_STOP_NONE  0
#define PYITD_STOP_KNOTS 1    /* baseline has < min_extrema extrema, ITD.py:404-416 */
#define PYITD_STOP_ITER  2    /* counter > max_iteration ("Out of time!"), ITD.py:418-426 */

/* call-level error codes */
#define PYITD_E_INVALID  (-1) /* bad argument */
#define PYITD_E_CUDA     (-2) /* CUDA runtime error, see pyitd_last_error() */
#define PYITD_E_NOMEM    (-3) /* workspace allocation failed */
#define PYITD_E_NODEVICE (-4) /* no usable sm_100 device */

/* option flags for pyitd_plan_create */
#define PYITD_OPT_BASELINES  1  /* also produce the per-level baselines (ITD.get_baselines, ITD.py:436) */
#define PYITD_OPT_ZERO_TAIL  2  /* zero-fill rows >= n_rows[signal] (the reference's zeros((22,N)) look) */

typedef struct pyitd_plan pyitd_plan;

PYITD_API int         pyitd_abi_version(void);
PYITD_API const char *pyitd_last_error(void);      /* thread-local text of the last failure */
PYITD_API int         pyitd_device_count(void);

/*
 * A plan owns the device workspace for one problem shape (like an FFT plan): the ping-pong carry
 * buffers that keep every signal resident in HBM between levels, the compacted knot tables and
 * the look-back descriptors.  Replaces the per-call numpy.zeros((22, N)) pair of ITD.py:384-387.
 */
PYITD_API int  pyitd_plan_create(pyitd_plan **plan, int device, int64_t n_signals, int64_t n_samples,
                       int dtype, int max_iteration, int min_extrema, int options);
PYITD_API void pyitd_plan_destroy(pyitd_plan *plan);
PYITD_API int     pyitd_plan_rows(const pyitd_plan *plan);             /* max_iteration + 2 */
PYITD_API int64_t pyitd_plan_workspace_bytes(const pyitd_plan *plan);
PYITD_API int     pyitd_plan_launches(const pyitd_plan *plan);         /* kernel launches of the last call */
/* Optional code paths of this build: "sweep_fused_pairs" (always), "sweep_fused_scan" (the scan-free extraction 0 experiment,
 * only with -DPYITD_SWEEP_WITH_FUSED_SCAN).  1 if compiled in, else 0. */
PYITD_API int     pyitd_has_feature(const char *name);
/* Sweep path, last call (synchronises with the device): how many PAIRS of extractions ran as one fused item (the baseline
 * between them never stored: one pass over X_e, 32 bytes per sample instead of 48 for ITD.py:79-121 twice; the knots of
 * that baseline are predicted from the knot table and the pass checks the prediction on every sample), how many pairs were
 * not tried after the prediction (too few or too many knots), and how many fused passes failed their check (the extraction
 * was then redone on its own; that signal tries no further pairs).  PYITD_SWEEP_FUSE=0 at plan creation turns pairs off. */
PYITD_API int     pyitd_plan_sweep_stats(pyitd_plan *plan, int64_t *fused_pairs, int64_t *pairs_skipped, int64_t *pairs_failed);
/* Which kernel family pyitd_decompose_* uses for this shape: PYITD_PATH_RESIDENT (signal kept on chip by a
 * thread-block cluster, one launch per batch), PYITD_PATH_STREAM (one CTA per signal, carry in HBM) or
 * PYITD_PATH_LOOKBACK (one CTA per tile with a look-back chain, carry in HBM: a handful of signals) or
 * PYITD_PATH_STRIDED (ONE long signal: persistent CTAs stride over its tiles with the streaming pipeline).  *cluster_size (may be NULL)
 * receives the CTAs per cluster of the resident kernel, else 1. */
#define PYITD_PATH_LOOKBACK 0
#define PYITD_PATH_STREAM   1
#define PYITD_PATH_RESIDENT 2
#define PYITD_PATH_STRIDED  3
#define PYITD_PATH_SWEEP    4   /* many signals: ONE persistent launch decomposes the whole batch, every level (ticket-
                                   scheduled (level, signal) items, one CTA per item, carry in HBM) */
#define PYITD_PATH_COOP     5   /* up to 16 signals shorter than 2^18 samples (the reference's own use, ITD.py:500-503: ONE signal):
                                   ONE cooperative launch, each signal kept in the shared memory of a group of CTAs, one group
                                   barrier per extraction */
PYITD_API int     pyitd_plan_path(const pyitd_plan *plan, int *cluster_size);

/* Stream path only: cut the batch into `groups` contiguous signal ranges (1..16), each with its own launch
 * chain on an internal stream forked from / joined to the caller's stream, so the last partial wave of one
 * range's launch overlaps another range's next launch.  Results do not depend on it.  The default is 2 for
 * batches of 1024 signals or more, else 1; PYITD_GROUPS in the environment overrides it at plan creation. */
PYITD_API int     pyitd_plan_set_groups(pyitd_plan *plan, int groups);
PYITD_API int     pyitd_plan_groups(const pyitd_plan *plan);

/* Measurement aid: with timing enabled every kernel launch of pyitd_decompose_device is bracketed by
 * CUDA events on the launching stream; pyitd_plan_launch_times waits for the last one and returns the
 * number of launches, writing their durations in ms (launch 0 = knot scan, 1.. = one per level).  With
 * more than one launch group the chains overlap, and ONE duration is returned: the whole call, fork to join. */
PYITD_API int     pyitd_plan_enable_timing(pyitd_plan *plan, int enable);
PYITD_API int     pyitd_plan_launch_times(pyitd_plan *plan, float *ms, int capacity);

/* Measurement aid: a plain streaming kernel with the level kernel's traffic mix -- reads n_doubles float64 from x,
 * writes n_doubles to y and to z (128-bit accesses, `ctas` blocks of 256 threads, grid-stride).  Its GB/s is the
 * memory system's practical ceiling for a 1 : 2 read : write stream, next to the copy peak in MEASURED_PEAKS.json.
 * chunk_doubles == 0: grid-stride; > 0: every block streams its own contiguous ranges of that many doubles (the
 * access pattern of one CTA per signal). */
PYITD_API int     pyitd_probe_mixed_traffic(const void *x, void *y, void *z, int64_t n_doubles, int ctas,
                                            int64_t chunk_doubles, void *stream);

/*
 * Replaces ITD.itd(data, max_iteration) (ITD.py:351-433) for a batch of independent signals.
 *   x           [n_signals, n_samples]            dtype per plan (device memory)
 *   rotations   [n_signals, rows, n_samples]      rows [0, n_rows[s]) = proper rotations, last = trend
 *   baselines   [n_signals, rows, n_samples]      NULL unless PYITD_OPT_BASELINES; valid rows:
 *                                                 n_rows-1 on the knot stop (ITD.py:414), n_rows on the
 *                                                 iteration stop with a zero last row (ITD.py:424)
 *   n_rows      [n_signals] int32
 *   knot_counts [n_signals, rows] int32           extrema of each new baseline = what ITD.py:403 prints
 *   input_knots [n_signals] int32 (may be NULL)   interior knots of the input itself
 *   stop_kind   [n_signals] int32 (may be NULL)
 *   status      [n_signals] int32
 * No host synchronisation happens inside; the stop test runs on the device.
 */
PYITD_API int pyitd_decompose_device(pyitd_plan *plan, const void *x, void *rotations, void *baselines,
                           int32_t *n_rows, int32_t *knot_counts, int32_t *input_knots,
                           int32_t *stop_kind, int32_t *status, void *stream);

/* Same call with HOST buffers: copies x in, runs, copies every output back, synchronises.  This is
 * the entry a non-CUDA host (the reference's numpy caller) binds. */
PYITD_API int pyitd_decompose_host(pyitd_plan *plan, const void *x, void *rotations, void *baselines,
                         int32_t *n_rows, int32_t *knot_counts, int32_t *input_knots,
                         int32_t *stop_kind, int32_t *status);

/*
 * Replaces itd_baseline_extract(data) -> (rotation, baseline) (ITD.py:79-121): one sifting level.
 *   rotation, baseline [n_signals, n_samples]; knot_count [n_signals] = interior knots of x.
 */
PYITD_API int pyitd_extract_level_device(pyitd_plan *plan, const void *x, void *rotation, void *baseline,
                               int32_t *knot_count, int32_t *status, void *stream);

/*
 * one sifting level with SUPPLIED knots -- detect the knots once (pyitd_find_knots_device on a
 * reference channel, or any ascending list), then apply the knot baseline + interpolation of ITD.py:95-119 to
 * other channels or to updated data.  This is the "retention and reuse of extrema ... along multiple channels"
 * mode of the reference's C++ port (itd.cpp:41-44; compute_extrema == false at itd.cpp:156-169), on ITD.py's own
 * interpolant.  With a signal's own knots it equals pyitd_extract_level_device bit for bit.
 *   knots       [n_knot_rows, knot_capacity] int32, ascending interior indices in [1, n_samples - 2]
 *   knot_count  [n_knot_rows] int32
 *   n_knot_rows n_signals (one list per signal) or 1 (one list shared by every signal)
 *   status      PYITD_ST_BAD_KNOTS when a list is not strictly increasing inside [1, n-2] (that signal is then
 *               processed with the empty list), PYITD_ST_ZERO_DX / PYITD_ST_NONFINITE as elsewhere
 */
#define PYITD_ST_BAD_KNOTS 8
PYITD_API int pyitd_extract_with_knots_device(pyitd_plan *plan, const void *x, const int32_t *knots,
                                    int64_t knot_capacity, const int32_t *knot_count, int64_t n_knot_rows,
                                    void *rotation, void *baseline, int32_t *status, void *stream);

/*
 * one level of the cubic-spline baseline variant -- replaces
 * itd_baseline_extract(data) -> (rotation, baseline) of MEITD.py:303-338 and
 * itd_baseline_extract_modified(x) -> baseline of numba_accelerated_itd.py:183-211.  Knots and knot baseline as in
 * ITD.py, end knots from the odd-reflected pad (MEITD.py:323-325), baseline = the interpolating cubic spline with
 * not-a-knot ends through (tau_k, L_k) (scipy splrep(k=3) + splev in the reference, MEITD.py:330-333), evaluated at
 * every sample including the last; rotation = x - baseline (MEITD.py:335).  PYITD_F64 and PYITD_F32_MIXED plans.
 *   rotation   [n_signals, n_samples] or NULL (the numba variant returns the baseline only)
 *   baseline   [n_signals, n_samples]
 *   knot_count [n_signals] int32 = interior knots of x
 *   min_knots  signals with fewer interior knots get baseline = x, rotation = 0 (10 in numba_accelerated_itd.py:
 *              188-191; values below 2 act as 2)
 *   status     PYITD_ST_FEW_KNOTS when a signal has fewer than 2 interior knots (the reference's splrep raises
 *              TypeError "m > k must hold"); PYITD_ST_NONFINITE as elsewhere
 */
#define PYITD_ST_FEW_KNOTS 16
PYITD_API int pyitd_extract_spline_device(pyitd_plan *plan, const void *x, void *rotation, void *baseline,
                                int32_t *knot_count, int32_t *status, int min_knots, void *stream);

/*
 * the 2-D "crossways" ensemble ITD of siftED2D.ipynb (code cell 1, raw JSON :233-278).
 *
 * pyitd_crossways_device replaces crossways_itd_baseline_extract(data) for a batch of images: the spline baseline
 * (pyitd_extract_spline_device, min_knots = 10 in the notebook) along every row, along every column, then along the
 * rows of the column result and the columns of the row result; out = (lengthwise + crosswise) / 2.
 *   row_plan   plan for (n_images * height) signals of width samples
 *   col_plan   plan for (n_images * width) signals of height samples   (same dtype and device)
 *   images,out [n_images, height, width] C-contiguous, plan dtype
 *   scratch    pyitd_crossways_scratch_bytes(...) bytes of device memory
 *
 * pyitd_ensemble2d_device replaces retrieve_statistical_image_component(data) with the noise draws SUPPLIED
 * (the notebook draws them from numba's unseeded generator): members data + v_e and data - v_e for e < draws,
 * crossways on each, lowpass = mean over e of the pair means.  totalextract2d = [data - lowpass, lowpass].
 *   image [height, width]; noise [draws, height, width]; lowpass [height, width]
 *   row_plan / col_plan sized for n_images = 2 * draws; scratch pyitd_ensemble2d_scratch_bytes(...) bytes
 * Non-finite input is not supported (no status is reported here).
 */
PYITD_API int64_t pyitd_crossways_scratch_bytes(const pyitd_plan *row_plan, int64_t n_images, int64_t height, int64_t width);
PYITD_API int pyitd_crossways_device(pyitd_plan *row_plan, pyitd_plan *col_plan, const void *images, void *out, void *scratch,
                           int64_t n_images, int64_t height, int64_t width, int min_knots, void *stream);
PYITD_API int64_t pyitd_ensemble2d_scratch_bytes(const pyitd_plan *row_plan, int64_t draws, int64_t height, int64_t width);
PYITD_API int pyitd_ensemble2d_device(pyitd_plan *row_plan, pyitd_plan *col_plan, const void *image, const void *noise,
                            void *lowpass, void *scratch, int64_t draws, int64_t height, int64_t width, int min_knots,
                            void *stream);

/*
 * post-decomposition analytics on rows that are still in device memory (no plan needed).
 *
 * pyitd_wpe_device replaces weighted_permutation_entropy(time_series, order=3, normalize) (MEITD.py:79-128), which the
 * reference evaluates per rotation (MEITD.py:346, :374, :547), for every row of a matrix.
 *   rows        [n_rows_total, n_samples], dtype PYITD_F64 (double) or any other value (float)
 *   valid_rows  NULL, or [n_rows_total / rows_per_signal] int32 (e.g. n_rows of pyitd_decompose_device): rows whose
 *               index inside their signal is >= valid_rows[signal] get NaN
 *   out         [n_rows_total] float64
 * The reference accumulates the weighted counts sequentially in float64; the kernel reduces them in parallel with a
 * double-double combine, so results agree to ~1e-14 relative, not bit for bit.
 *
 * pyitd_column_fsum_device replaces shewchuk(a) (helperfunctions.py:2-9) and shewchuk_sum (ITD.py:475-481):
 * column_sums[s, t] = math.fsum(rows[s, :valid, t]) -- exactly rounded, bit-identical to CPython's fsum -- and
 * totals[s] (may be NULL) = their sum over t in double-double arithmetic (within 1 ulp of math.fsum).
 *   rows [n_signals, rows_per_signal, n_samples]; rows_per_signal <= 64; n_signals <= 65535
 */
PYITD_API int pyitd_wpe_device(const void *rows, int64_t n_rows_total, int64_t n_samples, int dtype, int order,
                     int normalize, const int32_t *valid_rows, int64_t rows_per_signal, double *out, void *stream);
PYITD_API int pyitd_column_fsum_device(const void *rows, int64_t n_signals, int64_t rows_per_signal, int64_t n_samples,
                             int dtype, const int32_t *valid_rows, double *column_sums, double *totals, void *stream);

/*
 * Replaces detect_peaks (ITD.py:33-76) and the knot merge around it (ITD.py:87-88, :97).
 *   kinds      PYITD_KNOTS_VALLEYS = detect_peaks(x), PYITD_KNOTS_PEAKS = detect_peaks(-x),
 *              PYITD_KNOTS_BOTH = sort(unique(hstack(both))) = the knot set of one level
 *   knots      [n_signals, knot_capacity] int32, ascending, first knot_count[s] entries valid
 *   knot_count [n_signals] int32 (the true count even when it exceeds knot_capacity)
 */
#define PYITD_KNOTS_VALLEYS 1
#define PYITD_KNOTS_PEAKS   2
#define PYITD_KNOTS_BOTH    3
PYITD_API int pyitd_find_knots_device(pyitd_plan *plan, const void *x, int kinds, int32_t *knots,
                            int64_t knot_capacity, int32_t *knot_count, int32_t *status, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PYITD_B200_H */
