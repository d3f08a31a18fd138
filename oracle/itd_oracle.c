/*
 * itd_oracle.c -- CPU restatement of the PyITD sifting loop.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker, not the product: only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  Nothing under pyitd_b200/
 * imports, links or executes it; the product path fails loudly without its CUDA library.
 *
 * Parity status: PINNED.  tests/golden/ holds outputs of the reference's own numba
 * implementation (imported from /root/reference/ITD.py by tests/golden/make_golden.py in the
 * build container) for the reference's only golden vector (PyITD.ipynb cell 2, 8000 samples),
 * the ITD.py:491-495 demo signal, config 1 and the edge cases; tests/test_oracle.py checks this
 * file against them bit for bit.
 *
 * Every function cites the reference lines it restates (paths relative to /root/reference).
 * The arithmetic is scalar IEEE-754 in the reference's operation order; build with
 * -ffp-contract=off so gcc never fuses a multiply-add (numba does not either without fastmath).
 *
 * Status codes (shared with include/pyitd_b200.h):
 *   0 ok, 1 zero delta-X in a segment (reference raises ZeroDivisionError, ITD.py:116),
 *   2 non-finite input (reference NaN path ITD.py:46-51,64-68 is not supported),
 *   3 signal shorter than 3 samples (undefined in the reference, ITD.py:42-43),
 *   8 a supplied knot list is not strictly increasing inside [1, n-2] (extract_with_knots only).
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <sched.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define ITD_OK 0
#define ITD_ZERO_DX 1
#define ITD_NONFINITE 2
#define ITD_TOO_SHORT 3

/* ---- extrema detection ------------------------------------------------------------------
 * ITD.py:33-76 detect_peaks(x) returns i with dx[i] > 0 and dx[i-1] <= 0 (a valley, plateau
 * resolved to its right-most sample); ITD.py:87-88 calls it on x and on -x and ITD.py:97 takes
 * the sorted union; ITD.py:70-73 drop index 0 and n-1.  dx is formed by subtraction exactly as
 * ITD.py:44 does.  Returns the number of interior knots K and writes them ascending to idx. */
int64_t itd_oracle_find_knots_f64(const double *x, int64_t n, int64_t *idx)
{
    int64_t k = 0;
    if (n < 3) return 0;
    double dprev = x[1] - x[0];
    for (int64_t i = 1; i <= n - 2; ++i) {
        double dnext = x[i + 1] - x[i];
        int valley = (dnext > 0.0) && (dprev <= 0.0);       /* detect_peaks(x)   */
        int peak = ((-dnext) > 0.0) && ((-dprev) <= 0.0);   /* detect_peaks(-x)  */
        if (valley || peak) idx[k++] = i;
        dprev = dnext;
    }
    return k;
}

int64_t itd_oracle_find_knots_f32(const float *x, int64_t n, int64_t *idx)
{
    int64_t k = 0;
    if (n < 3) return 0;
    float dprev = x[1] - x[0];
    for (int64_t i = 1; i <= n - 2; ++i) {
        float dnext = x[i + 1] - x[i];
        int valley = (dnext > 0.0f) && (dprev <= 0.0f);
        int peak = ((-dnext) > 0.0f) && ((-dprev) <= 0.0f);
        if (valley || peak) idx[k++] = i;
        dprev = dnext;
    }
    return k;
}

/* ---- one sifting level ------------------------------------------------------------------
 * ITD.py:95-119 with the knot list GIVEN: tau[1..K] hold the interior knots (ascending, inside
 * [1, n-2]); this function adds tau[0] = 0 and tau[K+1] = n-1 (ITD.py:98) and evaluates the knot
 * baseline, the baseline and the rotation.  With the knots of x itself it is the body of
 * itd_baseline_extract; with another channel's knots it is the "retain and reuse the extrema ...
 * along multiple channels" mode of the reference's C++ port (itd.cpp:41-44, compute_extrema ==
 * false at itd.cpp:156-169) applied to ITD.py's own interpolant.  knotL must have room for K+2. */
#define ITD_BAD_KNOTS 8
int itd_oracle_extract_with_knots_f64(const double *x, int64_t n, int64_t *tau, int64_t K,
                                      double *R, double *B, double *knotL)
{
    if (n < 3) return ITD_TOO_SHORT;
    for (int64_t i = 0; i < n; ++i)
        if (!isfinite(x[i])) return ITD_NONFINITE;
    for (int64_t k = 1; k <= K; ++k)
        if (tau[k] < 1 || tau[k] > n - 2 || (k > 1 && tau[k] <= tau[k - 1])) return ITD_BAD_KNOTS;
    tau[0] = 0;
    tau[K + 1] = n - 1;

    /* ITD.py:100-102: numpy.mean of two samples = (0.0 + a + b) / 2 */
    knotL[0] = ((0.0 + x[0]) + x[1]) / 2.0;
    knotL[K + 1] = ((0.0 + x[n - 2]) + x[n - 1]) / 2.0;

    /* ITD.py:106-110 with alpha = 0.5 (ITD.py:85); int64 / int64 is a true float64 divide */
    for (int64_t k = 1; k <= K; ++k) {
        double w = (double)(tau[k] - tau[k - 1]) / (double)(tau[k + 1] - tau[k - 1]);
        double d = x[tau[k + 1]] - x[tau[k - 1]];
        double p = w * d;
        double q = x[tau[k - 1]] + p;
        knotL[k] = 0.5 * q + 0.5 * x[tau[k]];
    }

    /* ITD.py:112-117: baseline is affine in the SIGNAL VALUE inside each [tau_k, tau_k+1);
     * sample n-1 is never written and keeps the zeros_like value (ITD.py:112). */
    int status = ITD_OK;
    for (int64_t k = 0; k <= K; ++k) {
        double xk = x[tau[k]];
        double den = x[tau[k + 1]] - xk;
        if (den == 0.0) status = ITD_ZERO_DX;            /* numba raises ZeroDivisionError */
        double s = (knotL[k + 1] - knotL[k]) / den;
        for (int64_t t = tau[k]; t < tau[k + 1]; ++t) {
            double u = x[t] - xk;
            double v = s * u;
            B[t] = knotL[k] + v;
        }
    }
    B[n - 1] = 0.0;

    /* ITD.py:119 */
    for (int64_t t = 0; t < n; ++t) R[t] = x[t] - B[t];
    return status;
}

/* ITD.py:79-121 itd_baseline_extract.  tau must have room for n entries, knotL for n.
 * Writes rotation R and baseline B (both length n); *K_out = interior knot count of x. */
int itd_oracle_extract_level_f64(const double *x, int64_t n, double *R, double *B,
                                 int64_t *tau, double *knotL, int64_t *K_out)
{
    if (n < 3) return ITD_TOO_SHORT;
    for (int64_t i = 0; i < n; ++i)
        if (!isfinite(x[i])) return ITD_NONFINITE;
    /* ITD.py:87-98: the knots of x itself */
    int64_t K = itd_oracle_find_knots_f64(x, n, tau + 1);
    if (K_out) *K_out = K;
    return itd_oracle_extract_with_knots_f64(x, n, tau, K, R, B, knotL);
}

/* Same operation sequence carried out entirely in IEEE binary32 (the "pure fp32" variant the
 * product exposes as dtype='f32'; there is no such path in the reference, whose signatures are
 * float64 only, ITD.py:33,79).  The knot weight is formed from exact integers in double and
 * rounded once to float, so that index differences above 2^24 stay exact. */
int itd_oracle_extract_with_knots_f32(const float *x, int64_t n, int64_t *tau, int64_t K,
                                      float *R, float *B, float *knotL)
{
    if (n < 3) return ITD_TOO_SHORT;
    for (int64_t i = 0; i < n; ++i)
        if (!isfinite(x[i])) return ITD_NONFINITE;
    for (int64_t k = 1; k <= K; ++k)
        if (tau[k] < 1 || tau[k] > n - 2 || (k > 1 && tau[k] <= tau[k - 1])) return ITD_BAD_KNOTS;
    tau[0] = 0;
    tau[K + 1] = n - 1;
    knotL[0] = ((0.0f + x[0]) + x[1]) / 2.0f;
    knotL[K + 1] = ((0.0f + x[n - 2]) + x[n - 1]) / 2.0f;
    for (int64_t k = 1; k <= K; ++k) {
        float w = (float)((double)(tau[k] - tau[k - 1]) / (double)(tau[k + 1] - tau[k - 1]));
        float d = x[tau[k + 1]] - x[tau[k - 1]];
        float p = w * d;
        float q = x[tau[k - 1]] + p;
        knotL[k] = 0.5f * q + 0.5f * x[tau[k]];
    }
    int status = ITD_OK;
    for (int64_t k = 0; k <= K; ++k) {
        float xk = x[tau[k]];
        float den = x[tau[k + 1]] - xk;
        if (den == 0.0f) status = ITD_ZERO_DX;
        float s = (knotL[k + 1] - knotL[k]) / den;
        for (int64_t t = tau[k]; t < tau[k + 1]; ++t) {
            float u = x[t] - xk;
            float v = s * u;
            B[t] = knotL[k] + v;
        }
    }
    B[n - 1] = 0.0f;
    for (int64_t t = 0; t < n; ++t) R[t] = x[t] - B[t];
    return status;
}

int itd_oracle_extract_level_f32(const float *x, int64_t n, float *R, float *B,
                                 int64_t *tau, float *knotL, int64_t *K_out)
{
    if (n < 3) return ITD_TOO_SHORT;
    for (int64_t i = 0; i < n; ++i)
        if (!isfinite(x[i])) return ITD_NONFINITE;
    int64_t K = itd_oracle_find_knots_f32(x, n, tau + 1);
    if (K_out) *K_out = K;
    return itd_oracle_extract_with_knots_f32(x, n, tau, K, R, B, knotL);
}

/* ---- level loop -------------------------------------------------------------------------
 * ITD.py:351-433 ITD.itd, with the row cap generalised from the hard-coded 22 (ITD.py:384-385)
 * to max_iteration + 2 as ITD_numba.py:102-103 intends, and the stop threshold exposed as
 * min_extrema (the reference uses 2, ITD.py:404).
 *
 * rotations : (max_iteration + 2, n) row-major; rows [0, *n_rows) are valid, the rest zero.
 * baselines : NULL or same shape; rows [0, *n_baselines) valid.  On the iteration stop the last
 *             valid row is the reference's never-written zero row (ITD.py:424).
 * knot_counts[e] : extrema of the baseline produced by extraction e -- the numbers ITD.py:403
 *             prints, one per loop pass; length max_iteration + 2.
 * input_knots : interior knots of the input itself (not printed by the reference).
 * stop_kind : 1 knot-count stop (ITD.py:404-416), 2 iteration stop (ITD.py:418-426).        */
int itd_oracle_decompose_f64(const double *x, int64_t n, int max_iteration, int min_extrema,
                             double *rotations, double *baselines, int *n_rows,
                             int *n_baselines, int *knot_counts, int64_t *input_knots,
                             int *stop_kind)
{
    const int rmax = max_iteration + 2;
    memset(rotations, 0, sizeof(double) * (size_t)rmax * (size_t)n);
    if (baselines) memset(baselines, 0, sizeof(double) * (size_t)rmax * (size_t)n);
    for (int e = 0; e < rmax; ++e) knot_counts[e] = 0;
    *n_rows = 0;
    if (n_baselines) *n_baselines = 0;
    if (stop_kind) *stop_kind = 0;
    if (n < 3) return ITD_TOO_SHORT;

    double *cur = (double *)malloc(sizeof(double) * (size_t)n);   /* input of the extraction  */
    double *R = (double *)malloc(sizeof(double) * (size_t)n);
    double *B = (double *)malloc(sizeof(double) * (size_t)n);
    int64_t *tau = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 2));
    double *knotL = (double *)malloc(sizeof(double) * (size_t)(n + 2));
    int status = ITD_OK;
    memcpy(cur, x, sizeof(double) * (size_t)n);

    int64_t K = 0;
    status = itd_oracle_extract_level_f64(cur, n, R, B, tau, knotL, &K);      /* ITD.py:389 */
    if (input_knots) *input_knots = K;
    int c = 0;                                                                /* ITD.py:390 */
    while (status == ITD_OK) {
        int64_t ne = itd_oracle_find_knots_f64(B, n, tau);                    /* ITD.py:400-402 */
        knot_counts[c] = (int)ne;
        if (ne < min_extrema) {                                               /* ITD.py:404 */
            /* final row = baselines[c-1] = the input of the discarded extraction; for c == 0
             * the reference reads row -1 of a zero buffer (ITD.py:410) */
            if (c > 0) memcpy(rotations + (size_t)c * n, cur, sizeof(double) * (size_t)n);
            *n_rows = c + 1;
            if (n_baselines) *n_baselines = c;                                /* ITD.py:414 */
            if (stop_kind) *stop_kind = 1;
            break;
        } else if (c > max_iteration) {                                       /* ITD.py:418 */
            double *row = rotations + (size_t)c * n;
            for (int64_t t = 0; t < n; ++t) row[t] = R[t] + B[t];             /* ITD.py:420 */
            *n_rows = c + 1;
            if (n_baselines) *n_baselines = c + 1;                            /* ITD.py:424 */
            if (stop_kind) *stop_kind = 2;
            break;
        } else {                                                              /* ITD.py:428-432 */
            memcpy(rotations + (size_t)c * n, R, sizeof(double) * (size_t)n);
            if (baselines) memcpy(baselines + (size_t)c * n, B, sizeof(double) * (size_t)n);
            memcpy(cur, B, sizeof(double) * (size_t)n);
            status = itd_oracle_extract_level_f64(cur, n, R, B, tau, knotL, &K);
            c += 1;
        }
    }
    free(cur); free(R); free(B); free(tau); free(knotL);
    return status;
}

/* The product's mixed variant (fp32 in/out, fp64 carry and arithmetic) is by construction
 * float32(itd_oracle_decompose_f64(float64(x32))); the pure-fp32 variant is this function. */
int itd_oracle_decompose_f32(const float *x, int64_t n, int max_iteration, int min_extrema,
                             float *rotations, float *baselines, int *n_rows,
                             int *n_baselines, int *knot_counts, int64_t *input_knots,
                             int *stop_kind)
{
    const int rmax = max_iteration + 2;
    memset(rotations, 0, sizeof(float) * (size_t)rmax * (size_t)n);
    if (baselines) memset(baselines, 0, sizeof(float) * (size_t)rmax * (size_t)n);
    for (int e = 0; e < rmax; ++e) knot_counts[e] = 0;
    *n_rows = 0;
    if (n_baselines) *n_baselines = 0;
    if (stop_kind) *stop_kind = 0;
    if (n < 3) return ITD_TOO_SHORT;
    float *cur = (float *)malloc(sizeof(float) * (size_t)n);
    float *R = (float *)malloc(sizeof(float) * (size_t)n);
    float *B = (float *)malloc(sizeof(float) * (size_t)n);
    int64_t *tau = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 2));
    float *knotL = (float *)malloc(sizeof(float) * (size_t)(n + 2));
    memcpy(cur, x, sizeof(float) * (size_t)n);
    int64_t K = 0;
    int status = itd_oracle_extract_level_f32(cur, n, R, B, tau, knotL, &K);
    if (input_knots) *input_knots = K;
    int c = 0;
    while (status == ITD_OK) {
        int64_t ne = itd_oracle_find_knots_f32(B, n, tau);
        knot_counts[c] = (int)ne;
        if (ne < min_extrema) {
            if (c > 0) memcpy(rotations + (size_t)c * n, cur, sizeof(float) * (size_t)n);
            *n_rows = c + 1;
            if (n_baselines) *n_baselines = c;
            if (stop_kind) *stop_kind = 1;
            break;
        } else if (c > max_iteration) {
            float *row = rotations + (size_t)c * n;
            for (int64_t t = 0; t < n; ++t) row[t] = R[t] + B[t];
            *n_rows = c + 1;
            if (n_baselines) *n_baselines = c + 1;
            if (stop_kind) *stop_kind = 2;
            break;
        } else {
            memcpy(rotations + (size_t)c * n, R, sizeof(float) * (size_t)n);
            if (baselines) memcpy(baselines + (size_t)c * n, B, sizeof(float) * (size_t)n);
            memcpy(cur, B, sizeof(float) * (size_t)n);
            status = itd_oracle_extract_level_f32(cur, n, R, B, tau, knotL, &K);
            c += 1;
        }
    }
    free(cur); free(R); free(B); free(tau); free(knotL);
    return status;
}

/* ---- batch drivers (one whole channel per thread; channels are independent) ---------------
 * Used only as the CPU baseline timer and to check batched GPU output.  x is (nsig, n)
 * row-major, rotations (nsig, rmax, n); per-signal scalars are arrays of length nsig and
 * knot_counts is (nsig, rmax).  Plain pthreads pulling channel indices from an atomic counter
 * (this image has no libgomp).  Returns the number of signals with a non-zero status. */
typedef struct {
    const void *x;
    void *rotations, *baselines;
    int64_t nsig, n;
    int max_iteration, min_extrema, is_f32;
    int *n_rows, *knot_counts, *status;
    atomic_long next;
} batch_job;

static void *batch_worker(void *arg)
{
    batch_job *j = (batch_job *)arg;
    const int rmax = j->max_iteration + 2;
    for (;;) {
        int64_t b = (int64_t)atomic_fetch_add(&j->next, 1);
        if (b >= j->nsig) break;
        int nb = 0, kind = 0;
        int64_t ik = 0;
        size_t xo = (size_t)b * (size_t)j->n, ro = (size_t)b * (size_t)rmax * (size_t)j->n;
        if (j->is_f32)
            j->status[b] = itd_oracle_decompose_f32(
                (const float *)j->x + xo, j->n, j->max_iteration, j->min_extrema,
                (float *)j->rotations + ro, j->baselines ? (float *)j->baselines + ro : NULL,
                &j->n_rows[b], &nb, j->knot_counts + (size_t)b * rmax, &ik, &kind);
        else
            j->status[b] = itd_oracle_decompose_f64(
                (const double *)j->x + xo, j->n, j->max_iteration, j->min_extrema,
                (double *)j->rotations + ro, j->baselines ? (double *)j->baselines + ro : NULL,
                &j->n_rows[b], &nb, j->knot_counts + (size_t)b * rmax, &ik, &kind);
    }
    return NULL;
}

int itd_oracle_max_threads(void)
{
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) {
        int c = CPU_COUNT(&set);
        if (c > 0) return c;
    }
    long c = sysconf(_SC_NPROCESSORS_ONLN);
    return c > 0 ? (int)c : 1;
}

static int run_batch(batch_job *j, int nthreads)
{
    if (nthreads <= 0) nthreads = itd_oracle_max_threads();
    if (nthreads > 256) nthreads = 256;
    if ((int64_t)nthreads > j->nsig) nthreads = (int)(j->nsig > 0 ? j->nsig : 1);
    atomic_init(&j->next, 0);
    pthread_t th[256];
    int started = 0;
    for (int i = 1; i < nthreads; ++i)
        if (pthread_create(&th[started], NULL, batch_worker, j) == 0) started++;
    batch_worker(j);
    for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
    int bad = 0;
    for (int64_t b = 0; b < j->nsig; ++b) bad += (j->status[b] != ITD_OK);
    return bad;
}

int itd_oracle_decompose_batch_f64(const double *x, int64_t nsig, int64_t n, int max_iteration,
                                   int min_extrema, double *rotations, double *baselines,
                                   int *n_rows, int *knot_counts, int *status, int nthreads)
{
    batch_job j = {x, rotations, baselines, nsig, n, max_iteration, min_extrema, 0,
                   n_rows, knot_counts, status, 0};
    return run_batch(&j, nthreads);
}

int itd_oracle_decompose_batch_f32(const float *x, int64_t nsig, int64_t n, int max_iteration,
                                   int min_extrema, float *rotations, float *baselines,
                                   int *n_rows, int *knot_counts, int *status, int nthreads)
{
    batch_job j = {x, rotations, baselines, nsig, n, max_iteration, min_extrema, 1,
                   n_rows, knot_counts, status, 0};
    return run_batch(&j, nthreads);
}

/* ==== SURVEY.md 8f rank 2: the time-domain cubic-spline baseline variant =====================
 * numba_accelerated_itd.py:183-211 itd_baseline_extract_modified(x) -> baseline and
 * MEITD.py:303-338 itd_baseline_extract(data) -> (rotation, baseline): same knots (the stencil of
 * matlab_detect_peaks on x and -x, numba_accelerated_itd.py:17-59 == ITD.py:33-76), same knot
 * baseline loop (numba_accelerated_itd.py:168-178 == ITD.py:106-110), but
 *   - the end knots are the mean of the odd-reflected pad (numba_accelerated_itd.py:199-203,
 *     MEITD.py:323-325): L_0 = ((2 x[0] - x[1]) + x[0]) / 2, L_{K+1} = (x[n-1] + (2 x[n-1] - x[n-2])) / 2;
 *   - the baseline is the cubic spline through (tau_k, L_k), k = 0..K+1, evaluated at every sample
 *     0..n-1 (custom_splrep / numba_splev, numba_accelerated_itd.py:70-165, :207-209).
 * The spline itself lives in a THIRD-PARTY dependency: scipy.interpolate.splrep(x, y, k=3) with the
 * default s = 0 (no weights), i.e. FITPACK curfit/fpcurf (scipy 1.18.1 in this image; the reference's
 * environment.yml pins none): the interpolating cubic spline with not-a-knot end conditions (FITPACK
 * places no knot at x[1] and x[m-2]).  Its published definition is restated here through the
 * second-derivative ("moment") equations, Thomas-solved:
 *      mu_i M_{i-1} + 2 M_i + lam_i M_{i+1} = 6 [(y_{i+1}-y_i)/h_i - (y_i-y_{i-1})/h_{i-1}] / (h_{i-1}+h_i),
 *      not-a-knot: (M_1-M_0)/h_0 = (M_2-M_1)/h_1 and the mirror image at the right end.
 * FITPACK solves the same interpolation problem by Givens rotations on the B-spline collocation
 * matrix, so the two agree to rounding, not bit for bit: tests/test_oracle.py pins this function to
 * the reference's output (tests/golden/spline_*.npz) at 1e-12 relative L2 (measured 2e-16 .. 4e-15).
 * Needs K + 2 >= 4 points (splrep raises TypeError "m > k must hold" otherwise): status 16. */
#define ITD_FEW_KNOTS 16
int itd_oracle_spline_level_f64(const double *x, int64_t n, double *R, double *B, int64_t *K_out)
{
    if (n < 3) return ITD_TOO_SHORT;
    for (int64_t i = 0; i < n; ++i)
        if (!isfinite(x[i])) return ITD_NONFINITE;
    int64_t *tau = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 2));
    int64_t K = itd_oracle_find_knots_f64(x, n, tau + 1);
    if (K_out) *K_out = K;
    if (K < 2) { free(tau); return ITD_FEW_KNOTS; }
    tau[0] = 0;
    tau[K + 1] = n - 1;
    const int64_t P = K + 2;
    double *y = (double *)malloc(sizeof(double) * (size_t)P * 4);
    double *M = y + P, *cp = M + P, *dp = cp + P;
    y[0] = ((2.0 * x[0] - x[1]) + x[0]) / 2.0;
    y[K + 1] = (x[n - 1] + (2.0 * x[n - 1] - x[n - 2])) / 2.0;
    for (int64_t k = 1; k <= K; ++k) {
        double w = (double)(tau[k] - tau[k - 1]) / (double)(tau[k + 1] - tau[k - 1]);
        double d = x[tau[k + 1]] - x[tau[k - 1]];
        y[k] = 0.5 * (x[tau[k - 1]] + w * d) + 0.5 * x[tau[k]];
    }
    /* rows i = 1..K of the moment equations, the two end rows with M_0 / M_{K+1} eliminated */
    for (int64_t i = 1; i <= K; ++i) {
        double h0 = (double)(tau[i] - tau[i - 1]), h1 = (double)(tau[i + 1] - tau[i]);
        double a = h0 / (h0 + h1), c = h1 / (h0 + h1), b = 2.0;
        double d = 6.0 * ((y[i + 1] - y[i]) / h1 - (y[i] - y[i - 1]) / h0) / (h0 + h1);
        if (i == 1) { double r = h0 / h1; b += a * (1.0 + r); c -= a * r; a = 0.0; }
        if (i == K) { double r = h1 / h0; b += c * (1.0 + r); a -= c * r; c = 0.0; }
        if (i == 1) { cp[i] = c / b; dp[i] = d / b; }
        else { double m = b - a * cp[i - 1]; cp[i] = c / m; dp[i] = (d - a * dp[i - 1]) / m; }
    }
    M[K] = dp[K];
    for (int64_t i = K - 1; i >= 1; --i) M[i] = dp[i] - cp[i] * M[i + 1];
    {
        double r0 = (double)(tau[1] - tau[0]) / (double)(tau[2] - tau[1]);
        double rK = (double)(tau[K + 1] - tau[K]) / (double)(tau[K] - tau[K - 1]);
        M[0] = (1.0 + r0) * M[1] - r0 * M[2];
        M[K + 1] = (1.0 + rK) * M[K] - rK * M[K - 1];
    }
    for (int64_t j = 0; j <= K; ++j) {
        double h = (double)(tau[j + 1] - tau[j]);
        double c1 = (y[j + 1] - y[j]) / h - h * (2.0 * M[j] + M[j + 1]) / 6.0;
        double c2 = M[j] / 2.0, c3 = (M[j + 1] - M[j]) / (6.0 * h);
        int64_t t1 = (j == K) ? n : tau[j + 1];          /* the last segment includes sample n-1 */
        for (int64_t t = tau[j]; t < t1; ++t) {
            double u = (double)(t - tau[j]);
            B[t] = y[j] + u * (c1 + u * (c2 + u * c3));
        }
    }
    for (int64_t t = 0; t < n; ++t) R[t] = x[t] - B[t];  /* MEITD.py:335 */
    free(y); free(tau);
    return ITD_OK;
}

/* ==== SURVEY.md 8f rank 4: post-decomposition analytics ========================================
 * (a) weighted permutation entropy of order 3, MEITD.py:79-128 (used per rotation at MEITD.py:346,
 *     :374, :547): windows (x[i], x[i+1], x[i+2]); pattern = stable argsort of the window (numpy's
 *     quicksort is an insertion sort below 16 elements, MEITD.py:106); hash = sum idx_k 3^k (:108);
 *     weight = population variance of the window (:107) = (((d0^2 + d1^2) + d2^2) / 3 with
 *     d = x - ((a + b) + c) / 3; weighted counts are accumulated per pattern in window order (:113-117);
 *     p = counts / sum(counts) over the patterns that occur, in ascending hash order (:122);
 *     pe = -sum p log2 p (:123), divided by log2(3!) when normalised (:124-125).                 */
double itd_oracle_wpe3_f64(const double *x, int64_t n, int normalize)
{
    /* ascending hash order of the six permutations: (2,1,0)=5 (1,2,0)=7 (2,0,1)=11 (0,2,1)=15 (1,0,2)=19 (0,1,2)=21 */
    double wc[6] = {0, 0, 0, 0, 0, 0};
    int64_t cnt[6] = {0, 0, 0, 0, 0, 0};
    for (int64_t i = 0; i + 2 < n; ++i) {
        double a = x[i], b = x[i + 1], c = x[i + 2];
        int slot;
        /* stable argsort of (a, b, c): ties keep index order */
        if (a <= b) {
            if (b <= c) slot = 5;            /* (0,1,2) -> 21 */
            else if (a <= c) slot = 3;       /* (0,2,1) -> 15 */
            else slot = 2;                   /* (2,0,1) -> 11 */
        } else {
            if (a <= c) slot = 4;            /* (1,0,2) -> 19 */
            else if (b <= c) slot = 1;       /* (1,2,0) -> 7  */
            else slot = 0;                   /* (2,1,0) -> 5  */
        }
        double mean = ((a + b) + c) / 3.0;
        double d0 = a - mean, d1 = b - mean, d2 = c - mean;
        double w = ((d0 * d0 + d1 * d1) + d2 * d2) / 3.0;
        wc[slot] += w;
        cnt[slot] += 1;
    }
    double total = 0.0;
    int first = 1;
    for (int k = 0; k < 6; ++k)
        if (cnt[k]) { total = first ? wc[k] : total + wc[k]; first = 0; }
    double acc = 0.0;
    first = 1;
    for (int k = 0; k < 6; ++k)
        if (cnt[k]) {
            double p = wc[k] / total;
            double term = p * log2(p);
            acc = first ? term : acc + term;
            first = 0;
        }
    double pe = -acc;
    if (normalize) pe /= log2(6.0);
    return pe;
}

/* (b) exactly rounded column sums of the output rows, helperfunctions.py:2-9 shewchuk(a) == the inner loop of
 *     ITD.py:475-481 shewchuk_sum: s[t] = math.fsum(rows[:, t]).  math.fsum is restated (CPython's algorithm:
 *     a non-overlapping expansion grown by two-sums, rounded once with the half-even correction).         */
static double fsum_exact(const double *v, int64_t count, int64_t stride)
{
    double partials[64];
    int np_ = 0;
    for (int64_t r = 0; r < count; ++r) {
        double xx = v[r * stride];
        int i = 0;
        for (int j = 0; j < np_; ++j) {
            double y = partials[j];
            if (fabs(xx) < fabs(y)) { double t = xx; xx = y; y = t; }
            double hi = xx + y;
            double lo = y - (hi - xx);
            if (lo != 0.0) partials[i++] = lo;
            xx = hi;
        }
        partials[i] = xx;
        np_ = i + 1;
    }
    double hi = 0.0;
    if (np_ > 0) {
        int n2 = np_;
        hi = partials[--n2];
        double lo = 0.0;
        while (n2 > 0) {
            double xx = hi;
            double y = partials[--n2];
            hi = xx + y;
            double yr = hi - xx;
            lo = y - yr;
            if (lo != 0.0) break;
        }
        if (n2 > 0 && ((lo < 0.0 && partials[n2 - 1] < 0.0) || (lo > 0.0 && partials[n2 - 1] > 0.0))) {
            double y = lo * 2.0;
            double xx = hi + y;
            double yr = xx - hi;
            if (y == yr) hi = xx;
        }
    }
    return hi;
}

void itd_oracle_column_fsum_f64(const double *rows, int64_t n_rows, int64_t n, double *out)
{
    for (int64_t t = 0; t < n; ++t) out[t] = fsum_exact(rows + t, n_rows > 64 ? 64 : n_rows, n);
}
