// itd_spline.cuh -- SURVEY.md 8f rank 2: one level of the cubic-spline baseline variant of ITD.
//
// Reference (paths relative to /root/reference): itd_baseline_extract of MEITD.py:303-338 (rotation and
// baseline) and itd_baseline_extract_modified of numba_accelerated_itd.py:183-211 (baseline only).  Knots and
// the knot baseline L_k are those of ITD.py (same stencil, same Frei-Osorio formula); the end knots are the
// mean of the odd-reflected pad (MEITD.py:323-325); the baseline is the interpolating cubic spline with
// not-a-knot ends through (tau_k, L_k) -- scipy.interpolate.splrep(k=3, s=0) + splev in the reference
// (MEITD.py:330-333) -- evaluated at every sample.
//
// ONE kernel after the ordinary knot scan (which leaves tau / X_k / flag mask / per-tile prefix):
//
//   spline_level_kernel  walks a signal in WINDOWS of 896 knots.  Per window:
//     1. the knot slice (tau, X) plus 64 halo knots either side is staged in shared memory; one thread per knot
//        rebuilds L_k, 1/h_k and the segment slopes;
//     2. the tridiagonal moment system  mu_i M_{i-1} + 2 M_i + lam_i M_{i+1} = d_i  is solved by TRUNCATED
//        CYCLIC REDUCTION in shared memory (two plain levels, then parallel cyclic reduction on the quarter).  The rows are strictly diagonally dominant (off-diagonal
//        sum <= 1/2 of the diagonal), so after s reduction steps the remaining coupling to rows 2^s away is
//        below (1/2)^(2^s): at most six steps (reach 63 knots, residual coupling < 5.5e-20) decouple every row.
//        A window therefore depends on no other window: the solve is streaming and embarrassingly parallel in
//        the knot index, unlike the sequential Thomas sweep of the CPU restatement (or FITPACK's QR sweep);
//     3. the four power coefficients of the window's segments stay in shared memory and the samples of those
//        segments are evaluated straight from them: segment id = prefix popcount of the flag mask, Horner,
//        R = x - B, one coalesced load and two coalesced stores per sample.
//
// HBM traffic per sample: x read by the scan and by this kernel (2 s), R and B written (2 s); per knot 12 B of
// (tau, X_k) written by the scan and read here.  Spline coefficients never leave the chip.
// No tensor cores: nothing is a contraction.
#pragma once

#include "itd_kernels.cuh"

namespace pyitd {

constexpr int kStFewKnots = 16;
constexpr int kSplSlots = 1024;                       // rows of one PCR window held in shared memory
constexpr int kSplHalo = 64;                          // >= 1 + 2 + 4 + 8 + 16 + 32
constexpr int kSplUseful = kSplSlots - 2 * kSplHalo;  // rows a window is responsible for
constexpr int kSplThreads = 256;
constexpr int kSplSteps = 6;

struct SplineParams {
    const void *x;       // [S, N] input type
    void *rot;           // [S, N] output type or null
    void *bas;           // [S, N] output type
    KnotTable tab;       // of x, from the knot scan
    int *status;
    int n, tiles, tile;  // tile = samples per tbase entry
    int min_knots;       // signals with fewer interior knots keep B = x, R = 0 (numba_accelerated_itd.py:188-191)
    int parts;           // CTAs per signal
};

struct SplineSmem {
    // knot-indexed arrays: entry m is knot k = base - 2 + m of the window (base = unknown index of PCR slot 0)
    int ts[kSplSlots + 4];        // tau_k
    double xs[kSplSlots + 4];     // X_k, then the segment slope (L_{k+1} - L_k) / h_k, then the coefficient c1
    double ys[kSplSlots + 4];     // L_k (= coefficient c0)
    double ih[kSplSlots + 4];     // 1 / h_k, h_k = tau_{k+1} - tau_k
    // slot-indexed (slot q = entry m - 2): PCR rows with a unit diagonal; afterwards r = M, a = c2, c = c3
    // (stored at padded positions P(q) = q + q / 16 so that strides of 1, 2 and 4 rows are bank-conflict free)
    double a[kSplSlots + kSplSlots / 16], c[kSplSlots + kSplSlots / 16], r[kSplSlots + kSplSlots / 16];
};

__device__ __forceinline__ int P(int q) { return q + (q >> 4); }

// 1 / b to ~1 ulp: MUFU seed (about 20 bits) + two Newton steps.  The spline solve is held to 1e-9 against the
// reference's FITPACK solve, not to bit equality, so the ~40-instruction IEEE division is not needed here.
__device__ __forceinline__ double fast_rcp(double b) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(b));
    double e = __fma_rn(-b, x, 1.0);
    x = __fma_rn(x, e, x);
    e = __fma_rn(-b, x, 1.0);
    return __fma_rn(x, e, x);
}

template <typename InT, typename CarryT, typename OutT>
__global__ void __launch_bounds__(kSplThreads, 4) spline_level_kernel(const SplineParams p) {
    extern __shared__ __align__(16) unsigned char spl_smem_raw[];
    SplineSmem &sm = *reinterpret_cast<SplineSmem *>(spl_smem_raw);
    const int sig = blockIdx.x / p.parts, part = blockIdx.x % p.parts;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = p.tab.kcount[sig];
    const int n = p.n;
    const InT *x = reinterpret_cast<const InT *>(p.x) + (long long)sig * n;
    OutT *rot = p.rot ? reinterpret_cast<OutT *>(p.rot) + (long long)sig * n : nullptr;
    OutT *bas = reinterpret_cast<OutT *>(p.bas) + (long long)sig * n;

    if (K < 2 || K < p.min_knots) {
        // numba_accelerated_itd.py:188-191: too few extrema -> the input is returned as the baseline; with fewer than
        // two knots the reference's splrep raises (m > k must hold): reported through the status word
        if (part == 0 && tid == 0 && K < 2) atomicOr(p.status + sig, kStFewKnots);
        for (int t = part * kSplThreads + tid; t < n; t += p.parts * kSplThreads) {
            bas[t] = (OutT)__ldg(x + t);
            if (rot) rot[t] = (OutT)0;
        }
        return;
    }
    const int *tau = p.tab.tau + (long long)sig * p.tab.kstride;
    const CarryT *xk = reinterpret_cast<const CarryT *>(p.tab.xk) + (long long)sig * p.tab.kstride;
    const unsigned *mask = p.tab.mask + (long long)sig * p.tab.mstride;

    // end knots: mean of the odd-reflected pad (MEITD.py:323-325): pad = 2 * edge - neighbour
    const double x0 = (double)__ldg(x), x1 = (double)__ldg(x + 1);
    const double xn1 = (double)__ldg(x + n - 1), xn2 = (double)__ldg(x + n - 2);
    const double y_first = __dmul_rn(__dadd_rn(__dsub_rn(__dmul_rn(2.0, x0), x1), x0), 0.5);
    const double y_last = __dmul_rn(__dadd_rn(xn1, __dsub_rn(__dmul_rn(2.0, xn1), xn2)), 0.5);
    constexpr int M = kSplSlots + 4;

    constexpr int kPre = (M + kSplThreads - 1) / kSplThreads;      // knot-list entries per thread and window
    int tpre[kPre];
    bool have_pre = false;
#pragma unroll
    for (int i = 0; i < kPre; ++i) tpre[i] = 0;
    for (int win = part; win * kSplUseful < K; win += p.parts) {
        // PCR slot q holds unknown (interior knot) i = base + q = knot entry m = q + 2; rows outside [1, K] are
        // identity rows
        const int base = 1 + win * kSplUseful - kSplHalo;
        __syncthreads();                                   // previous window's readers are done
#pragma unroll
        for (int i = 0; i < kPre; ++i) {
            const int m = tid + i * kSplThreads;
            if (m < M) {
                const int k = base - 2 + m;
                const bool in = (k >= 0 && k <= K + 1);
                sm.ts[m] = have_pre ? tpre[i] : (in ? __ldg(tau + k) : 0);
                sm.xs[m] = in ? (double)__ldg(xk + k) : 0.0;
            }
        }
        {
            // the NEXT window's slice of the knot list while this window is solved and evaluated: tau into registers, X_k
            // towards L2.  The first use of the two loads above was the largest single stall of the kernel (a fifth of its warp
            // samples on the first level, ncu source page).
            const int kn = base - 2 + p.parts * kSplUseful;            // first entry of this CTA's next window
            have_pre = (win + p.parts) * kSplUseful < K;
            if (have_pre) {
#pragma unroll
                for (int i = 0; i < kPre; ++i) {
                    const int k = kn + tid + i * kSplThreads;
                    tpre[i] = (tid + i * kSplThreads < M && k >= 0 && k <= K + 1) ? __ldg(tau + k) : 0;
                }
                const int kq = kn + tid * 16;
                if (tid < 80 && kq >= 0 && kq <= K + 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(xk + kq));
            }
        }
        __syncthreads();
        for (int m = tid; m < M; m += kSplThreads) {
            const int k = base - 2 + m;
            double y = 0.0, ih = 0.0;
            if (k == 0) y = y_first;
            else if (k == K + 1) y = y_last;
            else if (k >= 1 && k <= K && m >= 1 && m <= M - 2) {
                // L_k of ITD.py:106-110 (the two outermost entries have no neighbour in the slice and are never used)
                const double w = (double)(sm.ts[m] - sm.ts[m - 1]) * fast_rcp((double)(sm.ts[m + 1] - sm.ts[m - 1]));
                y = 0.5 * (sm.xs[m - 1] + w * (sm.xs[m + 1] - sm.xs[m - 1])) + 0.5 * sm.xs[m];
            }
            if (k >= 0 && k <= K && m <= M - 2) ih = fast_rcp((double)(sm.ts[m + 1] - sm.ts[m]));
            sm.ys[m] = y;
            sm.ih[m] = ih;
        }
        __syncthreads();
        for (int m = tid; m < M - 1; m += kSplThreads)     // slopes overwrite X_k, which nobody reads any more
            sm.xs[m] = (sm.ys[m + 1] - sm.ys[m]) * sm.ih[m];
        __syncthreads();
        // rows of the moment equations  mu M_{i-1} + 2 M_i + lam M_{i+1} = 6 (s_i - s_{i-1}) / (h_{i-1} + h_i),
        // normalised to a unit diagonal; the end rows carry the not-a-knot conditions (M_0, M_{K+1} eliminated)
        for (int q = tid; q < kSplSlots; q += kSplThreads) {
            const int i = base + q, m = q + 2;
            double a = 0.0, c = 0.0, r = 0.0;
            if (i >= 1 && i <= K) {
                const double h0 = (double)(sm.ts[m] - sm.ts[m - 1]);       // tau_i - tau_{i-1}
                const double h1 = (double)(sm.ts[m + 1] - sm.ts[m]);       // tau_{i+1} - tau_i
                const double ihs = fast_rcp(h0 + h1);
                a = h0 * ihs;
                c = h1 * ihs;
                r = 6.0 * (sm.xs[m] - sm.xs[m - 1]) * ihs;
                double inv = 0.5;
                if (i == 1 || i == K) {
                    double b = 2.0;
                    if (i == 1) {
                        const double rho = h0 * sm.ih[m];
                        b += a * (1.0 + rho);
                        c -= a * rho;
                        a = 0.0;
                    }
                    if (i == K) {
                        const double rho = h1 * sm.ih[m - 1];
                        b += c * (1.0 + rho);
                        a -= c * rho;
                        c = 0.0;
                    }
                    inv = fast_rcp(b);
                }
                a *= inv;
                c *= inv;
                r *= inv;
            }
            sm.a[P(q)] = a;
            sm.c[P(q)] = c;
            sm.r[P(q)] = r;
        }
        __syncthreads();
        // Truncated cyclic reduction.  Two forward levels of plain cyclic reduction (rows 2i from their odd
        // neighbours, then rows 4i) leave 256 rows, one per thread; up to four steps of PARALLEL cyclic reduction
        // on those (distances 4, 8, 16, 32) decouple them -- it stops as soon as every remaining coupling is below
        // 1e-13, four orders of magnitude inside the 1e-9 parity tolerance (worst case (1/2)^64 after all six
        // levels) -- and two back-substitution levels return the odd rows.  1792 row updates per window instead of
        // the 6144 of six full PCR steps.  The rows live at padded positions P(q): conflict-free for strides 1, 2, 4.
        auto reduce_row = [&](int q, int d, double &na, double &nc, double &nr) {
            const double k1 = sm.a[P(q)], k2 = sm.c[P(q)];
            double am = 0.0, cm = 0.0, rm = 0.0, ap = 0.0, cp = 0.0, rp = 0.0;
            if (q - d >= 0) { am = sm.a[P(q - d)]; cm = sm.c[P(q - d)]; rm = sm.r[P(q - d)]; }
            if (q + d < kSplSlots) { ap = sm.a[P(q + d)]; cp = sm.c[P(q + d)]; rp = sm.r[P(q + d)]; }
            const double inv = fast_rcp(__fma_rn(-k1, cm, __fma_rn(-k2, ap, 1.0)));
            na = -(k1 * am) * inv;
            nc = -(k2 * cp) * inv;
            nr = __fma_rn(-k1, rm, __fma_rn(-k2, rp, sm.r[P(q)])) * inv;
        };
#pragma unroll
        for (int j = 0; j < 2; ++j) {                      // level 1: rows 2i (their odd neighbours are not written)
            const int q = 2 * (tid + j * kSplThreads);
            double na, nc, nr;
            reduce_row(q, 1, na, nc, nr);
            sm.a[P(q)] = na; sm.c[P(q)] = nc; sm.r[P(q)] = nr;
        }
        __syncthreads();
        {
            const int q = 4 * tid;                         // level 2: rows 4i from rows 4i +- 2
            double na, nc, nr;
            reduce_row(q, 2, na, nc, nr);
            sm.a[P(q)] = na; sm.c[P(q)] = nc; sm.r[P(q)] = nr;
            __syncthreads();
#pragma unroll 1
            for (int s = 0; s < 4; ++s) {                  // PCR among the rows 4i, in place through registers
                reduce_row(q, 4 << s, na, nc, nr);
                const int more = __syncthreads_or((fabs(na) > 1e-13) || (fabs(nc) > 1e-13));
                sm.a[P(q)] = na; sm.c[P(q)] = nc; sm.r[P(q)] = nr;
                __syncthreads();
                if (!more) break;
            }
            // back-substitution: rows 4i + 2 from the solved rows 4i, 4i + 4 (unit diagonal: M = r - a M- - c M+)
            const int q2 = q + 2;
            const double mp = (q2 + 2 < kSplSlots) ? sm.r[P(q2 + 2)] : 0.0;
            sm.r[P(q2)] = __fma_rn(-sm.a[P(q2)], sm.r[P(q2 - 2)], __fma_rn(-sm.c[P(q2)], mp, sm.r[P(q2)]));
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 2; ++j) {                      // odd rows from their solved even neighbours
            const int q = 2 * (tid + j * kSplThreads) + 1;
            const double mp = (q + 1 < kSplSlots) ? sm.r[P(q + 1)] : 0.0;
            sm.r[P(q)] = __fma_rn(-sm.a[P(q)], sm.r[P(q - 1)], __fma_rn(-sm.c[P(q)], mp, sm.r[P(q)]));
        }
        __syncthreads();
        // M_i = r[i] on the decoupled rows (slots [halo - 1, slots - halo] are exact to < 5.5e-20 relative).
        // Segments of this window: j = j_lo .. j_hi; the coefficients replace dead arrays in place
        // (c1 over the slope of the same segment, c2 / c3 over a / c of the same slot).
        const int j_lo = (win == 0) ? 0 : win * kSplUseful + 1;
        const int j_hi = min(K, win * kSplUseful + kSplUseful);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int q = tid + jj * kSplThreads;
            const int slot = kSplHalo - 1 + q;
            const int j = base + slot, m = slot + 2;
            if (q > kSplUseful || j < j_lo || j > j_hi) continue;
            const double h = (double)(sm.ts[m + 1] - sm.ts[m]);
            double Mj, Mj1;
            if (j == 0) {
                const double rho = h * sm.ih[m + 1];                       // h_0 / h_1
                Mj = (1.0 + rho) * sm.r[P(slot + 1)] - rho * sm.r[P(slot + 2)];
            } else {
                Mj = sm.r[P(slot)];
            }
            if (j == K) {
                const double rho = h * sm.ih[m - 1];                       // h_K / h_{K-1}
                Mj1 = (1.0 + rho) * sm.r[P(slot)] - rho * sm.r[P(slot - 1)];
            } else {
                Mj1 = sm.r[P(slot + 1)];
            }
            const double sixth = 1.0 / 6.0;
            sm.xs[m] = sm.xs[m] - h * (2.0 * Mj + Mj1) * sixth;            // only this thread touches xs[m], a / c[slot]
            sm.a[P(slot)] = 0.5 * Mj;
            sm.c[P(slot)] = (Mj1 - Mj) * sm.ih[m] * sixth;
        }
        __syncthreads();

        // evaluation of the window's samples: [tau_{j_lo}, tau_{j_hi + 1}), plus sample n-1 in the last window.
        // Four blocks of 256 samples per iteration (one sample per thread and block, one flag word per warp and
        // block).  Segment id of sample t = j_lo + flagged samples in (t_lo, t]: every warp keeps the running count
        // itself from one 128-byte load of the iteration's 32 flag words and a shuffle scan -- no shared memory,
        // no barrier, four independent x loads in flight per thread.
        const int t_lo = sm.ts[j_lo - (base - 2)];
        const int t_hi = (j_hi == K) ? n : sm.ts[j_hi + 1 - (base - 2)];
        constexpr int UN = 4;
        int run = j_lo;
        for (int tb = (t_lo & ~(UN * kSplThreads - 1)); tb < t_hi; tb += UN * kSplThreads) {
            double xv[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const int t = tb + u * kSplThreads + tid;
                xv[u] = (t >= t_lo && t < t_hi) ? (double)__ldg(x + t) : 0.0;
            }
            const int wi = (tb >> 5) + lane;               // lane l holds flag word l of the iteration
            unsigned word = (wi < p.tab.mstride) ? __ldg(mask + wi) : 0u;
            if (wi * 32 <= t_lo) word = (wi * 32 + 31 <= t_lo) ? 0u : (word & ~(0xffffffffu >> (31 - (t_lo & 31))));
            const int c = __popc(word);
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            const int excl = incl - c;
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const int src = u * (kSplThreads / 32) + warp;
                const unsigned mw = __shfl_sync(0xffffffffu, word, src);
                const int before = __shfl_sync(0xffffffffu, excl, src);
                const int t = tb + u * kSplThreads + tid;
                if (t < t_lo || t >= t_hi) continue;
                const int j = run + before + __popc(mw & (0xffffffffu >> (31 - lane)));
                const int m = j - (base - 2);
                const double uu = (double)(t - sm.ts[m]);
                const double b = sm.ys[m] + uu * (sm.xs[m] + uu * (sm.a[P(m - 2)] + uu * sm.c[P(m - 2)]));
                __stcs(bas + t, (OutT)b);
                if (rot) __stcs(rot + t, (OutT)(xv[u] - b));       // MEITD.py:335
            }
            run += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
}

}  // namespace pyitd
