// itd_analytics.cuh -- SURVEY.md 8f rank 4: post-decomposition analytics on the rows while they are still in HBM.
//
//   wpe3_kernel          weighted permutation entropy of order 3 per row (MEITD.py:79-128; the reference calls it
//                        per rotation for its MEITD / XITD selection, MEITD.py:346, :374, :547).  One CTA per row,
//                        one pass: pattern of (x[i], x[i+1], x[i+2]) by three compares (stable argsort, ties keep
//                        index order), weight = population variance of the window, six weighted counts per thread,
//                        double-double block reduction, entropy by one thread.  HBM: s bytes per sample.
//   column_fsum_kernel   exactly rounded sum over the rows of every sample column: math.fsum of
//                        helperfunctions.py:2-9 / ITD.py:475-481, CPython's algorithm (non-overlapping expansion grown
//                        by two-sums, one final rounding with the half-even correction), one thread per column.
//   dd_total_kernel      per-signal total of those column sums in double-double (106-bit) arithmetic.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace pyitd {

constexpr int kFsumMaxRows = 64;

struct dd {
    double hi, lo;
};
__device__ __forceinline__ dd dd_add_d(dd a, double b) {          // Knuth two-sum + renormalisation
    const double s = __dadd_rn(a.hi, b);
    const double bb = __dsub_rn(s, a.hi);
    const double e = __dadd_rn(__dsub_rn(a.hi, __dsub_rn(s, bb)), __dsub_rn(b, bb));
    const double lo = __dadd_rn(a.lo, e);
    const double hi = __dadd_rn(s, lo);
    return {hi, __dsub_rn(lo, __dsub_rn(hi, s))};
}
__device__ __forceinline__ dd dd_add(dd a, dd b) { return dd_add_d(dd_add_d(a, b.hi), b.lo); }
__device__ __forceinline__ dd dd_shfl_xor(dd v, int o) {
    return {__shfl_xor_sync(0xffffffffu, v.hi, o), __shfl_xor_sync(0xffffffffu, v.lo, o)};
}

// a / 3 correctly rounded without the division sequence (Markstein: y = RN(1/3), q = RN(a y), r = a - 3 q exactly by
// fma, RN(q + r y) is the correctly rounded quotient; exact for every a whose quotient is a normal number)
__device__ __forceinline__ double div3(double a) {
    const double y = 1.0 / 3.0;
    const double q = __dmul_rn(a, y);
    const double r = __fma_rn(-3.0, q, a);
    return __fma_rn(r, y, q);
}

// rows: [R, n] (InT); valid_rows (may be null): row r belongs to signal r / rows_per_signal and is skipped (NaN) when its
// index inside the signal is >= valid_rows[signal]
template <typename InT>
__global__ void __launch_bounds__(256) wpe3_kernel(const InT *__restrict__ rows, long long n, const int *__restrict__ valid_rows,
                                                   int rows_per_signal, int normalize, double *__restrict__ out) {
    __shared__ double s_hi[8][6], s_lo[8][6];
    __shared__ int s_cnt[8][6];
    const long long r = blockIdx.x;
    if (valid_rows && (int)(r % rows_per_signal) >= valid_rows[r / rows_per_signal]) {
        if (threadIdx.x == 0) out[r] = __longlong_as_double(0x7ff8000000000000ll);
        return;
    }
    const InT *x = rows + r * n;
    __shared__ double s_acc[6][256];
#pragma unroll
    for (int k = 0; k < 6; ++k) s_acc[k][threadIdx.x] = 0.0;
    unsigned seen = 0;                                     // bit k: pattern k occurred
    constexpr int UN = 4;                                  // windows in flight per thread (12 loads)
    auto window = [&](const double a, const double b, const double c) {
        // ascending hash order of MEITD.py:108: (2,1,0) (1,2,0) (2,0,1) (0,2,1) (1,0,2) (0,1,2):
        //   a <= b ? (b <= c ? 5 : (a <= c ? 3 : 2)) : (a <= c ? 4 : (b <= c ? 1 : 0))
        // without branches (a six-way divergent branch tree per window is what the compiler made of the ternaries):
        // three compare bits index a packed table
        const unsigned idx = ((a <= b) ? 4u : 0u) | ((b <= c) ? 2u : 0u) | ((a <= c) ? 1u : 0u);
        const int slot = (int)((0x55324140u >> (4u * idx)) & 7u);
        // population variance of the window (MEITD.py:112-113).  The result is compared at 1e-9 absolute, not bit for
        // bit (the reference's own summation order inside numpy.var is not specified), so the three squares are
        // accumulated with fused multiply-adds and the two divisions by 3 are multiplications: 11 fp64 operations per
        // window instead of 22.
        const double third = 1.0 / 3.0;
        const double mean = __dmul_rn(__dadd_rn(__dadd_rn(a, b), c), third);
        const double d0 = __dsub_rn(a, mean), d1 = __dsub_rn(b, mean), d2 = __dsub_rn(c, mean);
        const double w = __dmul_rn(__fma_rn(d2, d2, __fma_rn(d1, d1, __dmul_rn(d0, d0))), third);
        seen |= 1u << slot;
        // the thread's six weighted counts live in shared memory, indexed by the pattern: one load, one add, one store
        // per window instead of six selects and six additions
        double *ap = &s_acc[slot][threadIdx.x];
        *ap = __dadd_rn(*ap, w);
    };
    // 32-bit indices (the host entry point rejects rows of 2^31 samples or more); the main loop has no bounds checks, the
    // tail takes single windows
    const int nwin = (int)(n - 2), bd = (int)blockDim.x;
    int i0 = (int)threadIdx.x;
    for (; i0 + (UN - 1) * bd < nwin; i0 += UN * bd) {
        double a[UN], b[UN], c[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const InT *q = x + i0 + u * bd;
            a[u] = (double)__ldg(q);
            b[u] = (double)__ldg(q + 1);
            c[u] = (double)__ldg(q + 2);
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) window(a[u], b[u], c[u]);
    }
    for (; i0 < nwin; i0 += bd) window((double)__ldg(x + i0), (double)__ldg(x + i0 + 1), (double)__ldg(x + i0 + 2));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // the 256 per-thread sums of a pattern are added as a plain fp64 tree (a double-double tree here cost every thread 1500
    // fp64 instructions per row, a fifth of the kernel, to protect sums whose 256 terms each were accumulated in plain fp64)
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        double v = s_acc[k][threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
        const int c = __any_sync(0xffffffffu, (seen >> k) & 1u);
        if (lane == 0) {
            s_hi[warp][k] = v;
            s_lo[warp][k] = 0.0;
            s_cnt[warp][k] = c;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot_w[6];
        bool present[6];
        for (int k = 0; k < 6; ++k) {
            dd v = {0.0, 0.0};
            int c = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
                v = dd_add(v, dd{s_hi[w][k], s_lo[w][k]});
                c |= s_cnt[w][k];
            }
            tot_w[k] = v.hi;
            present[k] = c != 0;
        }
        // MEITD.py:121-125: p over the patterns that occur, ascending hash order; -sum p log2 p; / log2(3!)
        double total = 0.0, acc = 0.0;
        bool first = true;
        for (int k = 0; k < 6; ++k)
            if (present[k]) { total = first ? tot_w[k] : total + tot_w[k]; first = false; }
        first = true;
        for (int k = 0; k < 6; ++k)
            if (present[k]) {
                const double p = tot_w[k] / total;
                const double term = p * log2(p);
                acc = first ? term : acc + term;
                first = false;
            }
        double pe = -acc;
        if (normalize) pe /= log2(6.0);
        out[r] = pe;
    }
}

// rows: [S, R, n]; column sums over the first valid_rows[s] (or all R) rows -> sums[S, n]
template <typename InT>
__global__ void __launch_bounds__(256) column_fsum_kernel(const InT *__restrict__ rows, int R, long long n,
                                                          const int *__restrict__ valid_rows, double *__restrict__ sums) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long s = blockIdx.y;
    if (t >= n) return;
    int nr = valid_rows ? valid_rows[s] : R;
    nr = nr < 0 ? 0 : (nr > R ? R : nr);
    const InT *col = rows + s * R * n + t;
    double partials[kFsumMaxRows];
    int np = 0;
    for (int r0 = 0; r0 < nr; r0 += 8) {
        double v[8];                                       // eight independent loads in flight, then the expansion
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (r0 + u < nr) ? (double)__ldg(col + (long long)(r0 + u) * n) : 0.0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (r0 + u >= nr) break;
            double x = v[u];
            int i = 0;
            for (int j = 0; j < np; ++j) {
                double y = partials[j];
                if (fabs(x) < fabs(y)) { const double tmp = x; x = y; y = tmp; }
                const double hi = __dadd_rn(x, y);
                const double lo = __dsub_rn(y, __dsub_rn(hi, x));
                if (lo != 0.0) partials[i++] = lo;
                x = hi;
            }
            partials[i] = x;
            np = i + 1;
        }
    }
    double hi = 0.0;
    if (np > 0) {
        hi = partials[--np];
        double lo = 0.0;
        while (np > 0) {
            const double x = hi, y = partials[--np];
            hi = __dadd_rn(x, y);
            const double yr = __dsub_rn(hi, x);
            lo = __dsub_rn(y, yr);
            if (lo != 0.0) break;
        }
        if (np > 0 && ((lo < 0.0 && partials[np - 1] < 0.0) || (lo > 0.0 && partials[np - 1] > 0.0))) {
            const double y = __dmul_rn(lo, 2.0);
            const double x = __dadd_rn(hi, y);
            if (y == __dsub_rn(x, hi)) hi = x;
        }
    }
    sums[s * n + t] = hi;
}

// totals[s] = sum_t sums[s, t] in double-double
__global__ void __launch_bounds__(256) dd_total_kernel(const double *__restrict__ sums, long long n, double *__restrict__ totals) {
    __shared__ double s_hi[8], s_lo[8];
    const double *v = sums + (long long)blockIdx.x * n;
    dd acc = {0.0, 0.0};
    for (long long i = threadIdx.x; i < n; i += blockDim.x) acc = dd_add_d(acc, __ldg(v + i));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc = dd_add(acc, dd_shfl_xor(acc, o));
    if ((threadIdx.x & 31) == 0) {
        s_hi[threadIdx.x >> 5] = acc.hi;
        s_lo[threadIdx.x >> 5] = acc.lo;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        dd t = {0.0, 0.0};
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t = dd_add(t, dd{s_hi[w], s_lo[w]});
        totals[blockIdx.x] = t.hi;
    }
}

}  // namespace pyitd
