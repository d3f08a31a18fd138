// itd_kernels.cuh -- sm_100a kernels of the HBM-streaming ITD path.
//
// One sifting level of the reference (/root/reference/ITD.py:79-121) plus the stop test of its
// driver loop (ITD.py:400-404) is ONE pass over the signal:
//
//   knot_scan_kernel   (once per call)  3-point extrema stencil on the input, warp-ballot +
//                      decoupled look-back compaction of (tau_k, X_k) into the knot table.
//   level_kernel       (once per level) reads X_l once; rebuilds the tile's slice of the knot
//                      baseline L_k and the per-segment slope in shared memory (one thread per
//                      knot, ITD.py:106-110,116); evaluates B = L_k + s_k (x - X_k), R = x - B
//                      (ITD.py:115-119); writes R (output row) and B (next level's X); runs the
//                      extrema stencil on B while it is still on chip and compacts the NEXT
//                      level's knot table through the same look-back chain; the tile that closes
//                      a signal applies the stop rule on the device.
//
// Algorithmic HBM traffic per sample-level: read X (s) + write R (s) + write B (s) = 3 s bytes,
// plus 12..16 bytes per knot.  Nothing here is a contraction: no tensor cores, by design.
//
// fp64 arithmetic uses __d*_rn intrinsics in the reference's operation order so that no
// multiply-add is ever fused; results are bit-identical to the numba reference.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace pyitd {

constexpr int kStopOpen = 0x7f7f7f7f;   // stop_e value of a signal that is still decomposing
constexpr int kStopKnots = 1;
constexpr int kStopIter = 2;
constexpr int kStZeroDx = 1;
constexpr int kStNonFinite = 2;
constexpr int kStBadKnots = 8;

constexpr unsigned kOptBaselines = 1u;
constexpr unsigned kOptZeroTail = 2u;
constexpr unsigned kOptContigTiles = 8u;  // internal (strided level kernel): a block takes a contiguous run of tiles instead of every G-th
constexpr unsigned kOptGivenKnots = 4u;   // internal: segment ids from the stored flag mask, not from a stencil on x

// ---------------------------------------------------------------------------------------------
// unfused IEEE arithmetic in the carry type
// ---------------------------------------------------------------------------------------------
template <typename T> struct Arith;
template <> struct Arith<double> {
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    // ITD.py:108: int64 / int64 -> float64 true division
    static __device__ __forceinline__ double ratio(int a, int b) { return __ddiv_rn((double)a, (double)b); }
};
template <> struct Arith<float> {
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    // exact integers, one rounding: stays correct for index gaps above 2^24
    static __device__ __forceinline__ float ratio(int a, int b) { return (float)__ddiv_rn((double)a, (double)b); }
};

// ITD.py:59 on x and on -x (ITD.py:87-88), unioned (ITD.py:97): right-most sample of a plateau wins
template <typename T>
__device__ __forceinline__ bool is_knot(T a, T b, T c) {
    return (a >= b && b < c) || (a <= b && b > c);
}
// kinds bit 0: valleys = detect_peaks(x) (ITD.py:87); bit 1: peaks = detect_peaks(-x) (ITD.py:88)
template <typename T>
__device__ __forceinline__ bool is_knot_kind(T a, T b, T c, int kinds) {
    return ((kinds & 1) && a >= b && b < c) || ((kinds & 2) && a <= b && b > c);
}

// numpy.mean of two samples as numba evaluates it (ITD.py:101-102): ((0 + a) + b) / 2
template <typename T>
__device__ __forceinline__ T mean2(T a, T b) {
    return Arith<T>::div(Arith<T>::add(Arith<T>::add((T)0, a), b), (T)2);
}

// ---------------------------------------------------------------------------------------------
// decoupled look-back over the tiles of ONE signal
// descriptor = tag(30) | state(2) | count(32); the tag is a per-launch sequence number, so the
// array never needs clearing between levels.
// ---------------------------------------------------------------------------------------------
constexpr unsigned long long kAgg = 1ull, kIncl = 2ull;

__device__ __forceinline__ unsigned long long pack_desc(unsigned tag, unsigned long long state, int count) {
    return ((unsigned long long)tag << 34) | (state << 32) | (unsigned long long)(unsigned)count;
}
__device__ __forceinline__ unsigned long long ld_desc(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_desc(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Called by warp 0 only.  Returns the number of flagged samples in all earlier tiles of the signal.
__device__ __forceinline__ int lookback_exclusive(unsigned long long *desc, int tile, unsigned tag,
                                                  int total, int lane) {
    if (tile == 0) {
        if (lane == 0) st_desc(desc, pack_desc(tag, kIncl, total));
        return 0;
    }
    if (lane == 0) st_desc(desc + tile, pack_desc(tag, kAgg, total));
    int excl = 0;
    int pos = tile - 1;
    for (;;) {
        const int idx = pos - lane;
        unsigned long long d = pack_desc(tag, kIncl, 0);   // virtual tiles left of tile 0
        if (idx >= 0) {
            d = ld_desc(desc + idx);
            while ((unsigned)(d >> 34) != tag || ((d >> 32) & 3ull) == 0ull) {
                __nanosleep(32);
                d = ld_desc(desc + idx);
            }
        }
        const unsigned incl = __ballot_sync(0xffffffffu, ((d >> 32) & 3ull) == kIncl);
        const int first = __ffs(incl) - 1;                 // nearest tile with an inclusive prefix
        int v = (first < 0 || lane <= first) ? (int)(unsigned)(d & 0xffffffffull) : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        excl += v;
        if (first >= 0) break;
        pos -= 32;
    }
    if (lane == 0) st_desc(desc + tile, pack_desc(tag, kIncl, excl + total));
    return excl;
}

// Block-wide: from per-(round, warp) ballots to exclusive word prefixes + the look-back base.
// s_wpre[w] = flagged samples of this tile before word w; returns base (all threads), total via ref.
template <int THREADS, int ITEMS>
__device__ __forceinline__ int tile_scan(const unsigned (&bal)[ITEMS], int *s_wcnt, int *s_wpre,
                                         int *s_misc, unsigned long long *desc_sig, int tile,
                                         unsigned tag, int &total) {
    constexpr int WARPS = THREADS / 32;
    constexpr int WORDS = WARPS * ITEMS;
    constexpr int PER_LANE = (WORDS + 31) / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) s_wcnt[r * WARPS + warp] = __popc(bal[r]);
    }
    __syncthreads();
    if (warp == 0) {
        int c[PER_LANE];
        int mine = 0;
#pragma unroll
        for (int i = 0; i < PER_LANE; ++i) {
            const int w = lane * PER_LANE + i;
            c[i] = (w < WORDS) ? s_wcnt[w] : 0;
            mine += c[i];
        }
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        int run = incl - mine;
#pragma unroll
        for (int i = 0; i < PER_LANE; ++i) {
            const int w = lane * PER_LANE + i;
            if (w < WORDS) s_wpre[w] = run;
            run += c[i];
        }
        const int tot = __shfl_sync(0xffffffffu, incl, 31);
        const int base = lookback_exclusive(desc_sig, tile, tag, tot, lane);
        if (lane == 0) {
            s_misc[0] = base;
            s_misc[1] = tot;
        }
    }
    __syncthreads();
    total = s_misc[1];
    return s_misc[0];
}

// ---------------------------------------------------------------------------------------------
// parameter blocks (plain pointers; element types are fixed by the template arguments)
// ---------------------------------------------------------------------------------------------
struct KnotTable {
    int *tau;        // [S, kstride]  tau_0 = 0, interior knots, tau_{K+1} = N-1
    void *xk;        // [S, kstride]  X_k = x[tau_k]           (carry type)
    int *tbase;      // [S, tiles+1] interior knots before each tile; [tiles] = K
    int *kcount;     // [S]      K
    void *endl;      // [S, 2]   L_0 and L_{K+1} (ITD.py:101-102) (carry type)
    unsigned *mask;  // [S, mstride] knot flags, bit (t & 31) of word (t >> 5)
    long long kstride;   // entries per signal in tau/xk: N rounded up to 4, + 4 (16-byte aligned rows)
    long long mstride;   // words per signal in mask: ceil(N/32) rounded up to 4
    // strided path only (itd_strided.cuh), shared by both tables: knots per GROUP of 32 tiles, accumulated by the level /
    // scan kernel with one reduction per non-empty tile, and the groups' exclusive prefix (tile_prefix_kernel)
    int *gsum;           // [S, gstride]  zero between uses (the prefix kernel clears what it read)
    int *gbase;          // [S, gstride]
    long long gstride;
    int *stau;           // [S, kstride]  tile-local compaction: the knots of tile i at [i * T, i * T + count_i)
    void *sxk;           // [S, kstride]  (carry type)
};

struct ScanParams {
    const void *x;   // [S, N] input type
    KnotTable out;
    unsigned long long *desc;   // [S, tiles]
    unsigned tag;
    int *status;     // [S]
    int *input_knots;  // [S] or null
    int n;           // samples per signal
    int tiles;
    int kinds;       // 1 valleys, 2 peaks, 3 both (the knot set)
    int sig0 = 0;    // first signal of this launch (stream kernels: one launch per signal group)
};

struct LevelParams {
    const void *in;      // [S, N]: user input (level 0) or the carry written by the previous level
    void *carry_out;     // [S, N] carry type: B_e, the next level's X
    const void *fix_src; // [S, N] carry type: X_{e-1}, source of the knot-stop trend row
    void *rot;           // [S, rows, N] output type
    void *bas;           // [S, rows, N] output type or null
    long long out_sig_stride;   // elements between signals in rot/bas
    KnotTable cur, next;
    unsigned long long *desc;
    unsigned tag;
    int *stop_e, *stop_kind, *n_rows, *knot_counts, *status;
    int n, tiles;
    int e;               // extraction index of this launch
    int emax;            // last extraction allowed = max_iteration + 1
    int rows;            // emax + 1
    int min_extrema;
    unsigned opts;
    int sig0 = 0;        // first signal of this launch (stream kernels: one launch per signal group)
    // knot baseline table of knot_ls_kernel (stream / strided level kernels): ls[sig * (lscap + 4) + k] = {L_k, slope of
    // segment [k, k+1)} for the signals with at most lscap interior knots; null: every tile computes its own
    const void *ls = nullptr;
    int lscap = 0;
};

// tiles with at most kLsTile interior knots take L_k / slopes from the knot_ls_kernel table (one TMA slice per tile)
constexpr int kLsTile = 92;
constexpr int kLsStage = kLsTile + 4;

// ---------------------------------------------------------------------------------------------
// knot_ls_kernel: the Frei-Osorio knot baseline with ONE THREAD PER KNOT (ITD.py:100-110) and the slope of the segment
// that starts at the knot (ITD.py:116), for the signals whose table is sparse enough (K <= lscap) that the level kernel
// would otherwise recompute the same few knots in every warp that touches their segments.  Same operations in the same
// order as the level kernels' own evaluation: the results are bit-identical.
// grid (signals, y): block (s, y) takes knots y * 256 + tid, + gridDim.y * 256, ...
// ---------------------------------------------------------------------------------------------
template <typename CarryT>
struct alignas(2 * sizeof(CarryT)) KnotBaseSlope {
    CarryT L, s;
};
template <typename CarryT>
__global__ void __launch_bounds__(256) knot_ls_kernel(KnotTable cur, void *ls_out, int lscap, int sig0, int e,
                                                      const int *stop_e, int *status) {
    using A = Arith<CarryT>;
    const int sig = blockIdx.x + sig0;
    if (e > stop_e[sig]) return;
    const int K = cur.kcount[sig];
    if (K > lscap) return;
    const int *tau = cur.tau + (long long)sig * cur.kstride;
    const CarryT *xk = reinterpret_cast<const CarryT *>(cur.xk) + (long long)sig * cur.kstride;
    const CarryT *endl = reinterpret_cast<const CarryT *>(cur.endl) + 2ll * sig;
    KnotBaseSlope<CarryT> *out = reinterpret_cast<KnotBaseSlope<CarryT> *>(ls_out) + (long long)sig * (lscap + 4);
    auto knot_L = [&](const int k) -> CarryT {
        if (k == 0) return endl[0];
        if (k >= K + 1) return endl[1];
        const CarryT w = A::ratio(tau[k] - tau[k - 1], tau[k + 1] - tau[k - 1]);
        const CarryT d = A::sub(xk[k + 1], xk[k - 1]);
        const CarryT qq = A::add(xk[k - 1], A::mul(w, d));
        return A::add(A::mul((CarryT)0.5, qq), A::mul((CarryT)0.5, xk[k]));
    };
    bool zero_dx = false;
    for (int k = blockIdx.y * 256 + threadIdx.x; k <= K + 1; k += gridDim.y * 256) {
        const CarryT L = knot_L(k);
        CarryT sl = (CarryT)0;
        if (k <= K) {
            const CarryT den = A::sub(xk[k + 1], xk[k]);
            sl = A::div(A::sub(knot_L(k + 1), L), den);
            zero_dx |= (den == (CarryT)0);
        }
        out[k] = KnotBaseSlope<CarryT>{L, sl};
    }
    if (zero_dx) atomicOr(status + sig, kStZeroDx);
}


// ---------------------------------------------------------------------------------------------
// knot_scan_kernel: extrema detection + compaction on the raw input (ITD.py:87-98)
// ---------------------------------------------------------------------------------------------
template <typename InT, typename CarryT, int THREADS, int ITEMS>
__global__ void __launch_bounds__(THREADS) knot_scan_kernel(const ScanParams p) {
    constexpr int T = THREADS * ITEMS;
    constexpr int WORDS = T / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CarryT *xs = reinterpret_cast<CarryT *>(smem_raw);          // xs[j] = x[t0 - 1 + j], j < T + 3
    int *s_wcnt = reinterpret_cast<int *>(xs + (T + 4));
    int *s_wpre = s_wcnt + WORDS;
    int *s_misc = s_wpre + WORDS;

    const int sig = blockIdx.x / p.tiles, tile = blockIdx.x % p.tiles;
    const int n = p.n, t0 = tile * T;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const InT *x = reinterpret_cast<const InT *>(p.x) + (long long)sig * n;

    CarryT xr[ITEMS];
    bool bad = false;
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
        const int t = t0 + r * THREADS + tid;
        xr[r] = (t < n) ? (CarryT)__ldg(x + t) : (CarryT)0;
        bad |= !isfinite(xr[r]);
        xs[1 + r * THREADS + tid] = xr[r];
    }
    if (tid < 3) {
        const int t = (tid == 0) ? t0 - 1 : t0 + T + tid - 1;
        xs[(tid == 0) ? 0 : T + tid] = (t >= 0 && t < n) ? (CarryT)__ldg(x + t) : (CarryT)0;
    }
    if (__syncthreads_or(bad)) {
        if (tid == 0) atomicOr(p.status + sig, kStNonFinite);
    }

    unsigned bal[ITEMS];
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
        const int j = 1 + r * THREADS + tid, t = t0 + r * THREADS + tid;
        const bool f = (t >= 1) && (t <= n - 2) && is_knot_kind(xs[j - 1], xs[j], xs[j + 1], p.kinds);
        bal[r] = __ballot_sync(0xffffffffu, f);
    }
    int total;
    const int base = tile_scan<THREADS, ITEMS>(bal, s_wcnt, s_wpre, s_misc,
                                               p.desc + (long long)sig * p.tiles, tile, p.tag, total);

    int *tau = p.out.tau + (long long)sig * p.out.kstride;
    CarryT *xk = reinterpret_cast<CarryT *>(p.out.xk) + (long long)sig * p.out.kstride;
    unsigned *gmask = p.out.mask + (long long)sig * p.out.mstride + tile * (T / 32);
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
        if (lane == 0 && t0 + r * THREADS + warp * 32 < n) gmask[r * (THREADS / 32) + warp] = bal[r];
        if ((bal[r] >> lane) & 1u) {
            const int rank = base + s_wpre[r * (THREADS / 32) + warp] + __popc(bal[r] & ((1u << lane) - 1u));
            tau[1 + rank] = t0 + r * THREADS + tid;
            xk[1 + rank] = xr[r];
        }
    }
    if (tid == 0) {
        int *tb = p.out.tbase + (long long)sig * (p.tiles + 1);
        CarryT *endl = reinterpret_cast<CarryT *>(p.out.endl) + 2ll * sig;
        tb[tile] = base;
        if (tile == 0) {
            tau[0] = 0;
            xk[0] = xs[1];
            endl[0] = mean2<CarryT>(xs[1], xs[2]);
        }
        if (tile == p.tiles - 1) {
            const int K = base + total;
            const int jl = n - 1 - t0 + 1;                  // xs index of sample n-1
            tb[p.tiles] = K;
            p.out.kcount[sig] = K;
            tau[K + 1] = n - 1;
            xk[K + 1] = xs[jl];
            endl[1] = mean2<CarryT>(xs[jl - 1], xs[jl]);
            if (p.input_knots) p.input_knots[sig] = K;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// level_kernel: one sifting level, fused with the next level's knot detection and the stop test
// ---------------------------------------------------------------------------------------------
template <typename InT, typename CarryT, typename OutT, int THREADS, int ITEMS>
__global__ void __launch_bounds__(THREADS) level_kernel(const LevelParams p) {
    using A = Arith<CarryT>;
    constexpr int T = THREADS * ITEMS;
    constexpr int WARPS = THREADS / 32;
    constexpr int WORDS = T / 32;
    constexpr int KC = T + 8;                                   // knot slice capacity
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CarryT *xs = reinterpret_cast<CarryT *>(smem_raw);          // xs[j] = x[t0 - 1 + j], j < T + 3
    CarryT *kX = xs + (T + 4);
    CarryT *kL = kX + KC;
    CarryT *kS = kL + KC;
    int *ktau = reinterpret_cast<int *>(kS + KC);
    int *s_wcnt = ktau + KC;
    int *s_wpre = s_wcnt + WORDS;
    int *s_misc = s_wpre + WORDS;

    const int sig = blockIdx.x / p.tiles, tile = blockIdx.x % p.tiles;
    const int n = p.n, t0 = tile * T, e = p.e;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long row_off = (long long)sig * p.out_sig_stride;

    // ---- signals that already stopped: trend-row fix-up (ITD.py:410-411) or nothing -----------
    const int se = p.stop_e[sig];
    if (e > se) {
        OutT *rot = reinterpret_cast<OutT *>(p.rot) + row_off;
        OutT *bas = p.bas ? reinterpret_cast<OutT *>(p.bas) + row_off : nullptr;
        if (e == se + 1 && p.stop_kind[sig] == kStopKnots) {
            // the discarded extraction `se` wrote R_se into row se; the reference returns
            // baselines[se-1] there, i.e. the INPUT of that extraction (zeros when se == 0)
            const CarryT *src = reinterpret_cast<const CarryT *>(p.fix_src) + (long long)sig * n;
#pragma unroll
            for (int r = 0; r < ITEMS; ++r) {
                const int t = t0 + r * THREADS + tid;
                if (t < n) {
                    rot[(long long)se * n + t] = (se == 0) ? (OutT)0 : (OutT)src[t];
                    if (bas && (p.opts & kOptZeroTail)) bas[(long long)se * n + t] = (OutT)0;
                }
            }
        }
        if ((p.opts & kOptZeroTail) && e < p.rows) {
#pragma unroll
            for (int r = 0; r < ITEMS; ++r) {
                const int t = t0 + r * THREADS + tid;
                if (t < n) {
                    rot[(long long)e * n + t] = (OutT)0;
                    if (bas) bas[(long long)e * n + t] = (OutT)0;
                }
            }
        }
        return;
    }
    if (e > p.emax) return;          // trailing fix-up launch: nothing is active any more

    // ---- 1. stage the tile of X_e (+1 sample left, +2 right) ----------------------------------
    const InT *x = reinterpret_cast<const InT *>(p.in) + (long long)sig * n;
    CarryT xr[ITEMS];
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
        const int t = t0 + r * THREADS + tid;
        xr[r] = (t < n) ? (CarryT)__ldg(x + t) : (CarryT)0;
        xs[1 + r * THREADS + tid] = xr[r];
    }
    if (tid < 3) {
        const int t = (tid == 0) ? t0 - 1 : t0 + T + tid - 1;
        xs[(tid == 0) ? 0 : T + tid] = (t >= 0 && t < n) ? (CarryT)__ldg(x + t) : (CarryT)0;
    }

    // ---- 2. this tile's slice of the knot table ------------------------------------------------
    const int *tbase = p.cur.tbase + (long long)sig * (p.tiles + 1);
    const int kb = tbase[tile];                 // interior knots strictly before t0 = segment of t0-1
    const int cnt = tbase[tile + 1] - kb;       // interior knots inside the tile
    const int K = p.cur.kcount[sig];
    const int lo = max(kb - 1, 0);
    const int hi = min(kb + cnt + 3, K + 1);
    const int m = hi - lo + 1;                  // <= cnt + 5 <= KC
    {
        const int *gtau = p.cur.tau + (long long)sig * p.cur.kstride + lo;
        const CarryT *gxk = reinterpret_cast<const CarryT *>(p.cur.xk) + (long long)sig * p.cur.kstride + lo;
        for (int j = tid; j < m; j += THREADS) {
            ktau[j] = gtau[j];
            kX[j] = gxk[j];
        }
    }
    __syncthreads();

    // ---- 3. knot baseline, one thread per knot (ITD.py:100-110) --------------------------------
    {
        const CarryT *endl = reinterpret_cast<const CarryT *>(p.cur.endl) + 2ll * sig;
        for (int j = tid; j < m; j += THREADS) {
            const int k = lo + j;
            CarryT L = (CarryT)0;
            if (k == 0) {
                L = endl[0];
            } else if (k == K + 1) {
                L = endl[1];
            } else if (j >= 1 && j + 1 < m) {
                const CarryT w = A::ratio(ktau[j] - ktau[j - 1], ktau[j + 1] - ktau[j - 1]);
                const CarryT d = A::sub(kX[j + 1], kX[j - 1]);
                const CarryT q = A::add(kX[j - 1], A::mul(w, d));
                L = A::add(A::mul((CarryT)0.5, q), A::mul((CarryT)0.5, kX[j]));
            }
            kL[j] = L;
        }
    }
    __syncthreads();
    // per-segment slope (ITD.py:116); segments kb .. min(kb+cnt+1, K) are the ones evaluated here
    {
        bool zero_dx = false;
        const int seg_hi = min(kb + cnt + 1, K);
        for (int j = tid; j + 1 < m; j += THREADS) {
            const int k = lo + j;
            const CarryT den = A::sub(kX[j + 1], kX[j]);
            kS[j] = A::div(A::sub(kL[j + 1], kL[j]), den);
            zero_dx |= (k >= kb && k <= seg_hi && den == (CarryT)0);
        }
        if (zero_dx) atomicOr(p.status + sig, kStZeroDx);
    }

    // ---- 4. segment id of every sample = inclusive prefix count of the knot flags --------------
    unsigned bal[ITEMS];
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
        const int j = 1 + r * THREADS + tid, t = t0 + r * THREADS + tid;
        const bool f = (t >= 1) && (t <= n - 2) && is_knot(xs[j - 1], xs[j], xs[j + 1]);
        bal[r] = __ballot_sync(0xffffffffu, f);
    }
    if (p.opts & kOptGivenKnots) {
        // supplied knots (pyitd_extract_with_knots_device): the table's flag mask is the truth, x's own extrema are not
        const unsigned *gm = p.cur.mask + (long long)sig * p.cur.mstride + tile * (T / 32);
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) bal[r] = (t0 + r * THREADS + warp * 32 < n) ? gm[r * WARPS + warp] : 0u;
    }
    if (lane == 0) {
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) s_wcnt[r * WARPS + warp] = __popc(bal[r]);
    }
    __syncthreads();        // also publishes kS
    if (warp == 0) {
        constexpr int PER_LANE = (WORDS + 31) / 32;
        int c[PER_LANE], mine = 0;
#pragma unroll
        for (int i = 0; i < PER_LANE; ++i) {
            const int w = lane * PER_LANE + i;
            c[i] = (w < WORDS) ? s_wcnt[w] : 0;
            mine += c[i];
        }
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        int run = incl - mine;
#pragma unroll
        for (int i = 0; i < PER_LANE; ++i) {
            const int w = lane * PER_LANE + i;
            if (w < WORDS) s_wpre[w] = run;
            run += c[i];
        }
    }
    __syncthreads();

    // ---- 5. B = L_k + s_k (x - X_k), R = x - B (ITD.py:115-119); stream both out ---------------
    OutT *rot = reinterpret_cast<OutT *>(p.rot) + row_off + (long long)e * n;
    OutT *bas = p.bas ? reinterpret_cast<OutT *>(p.bas) + row_off + (long long)e * n : nullptr;
    CarryT *carry = reinterpret_cast<CarryT *>(p.carry_out) + (long long)sig * n;
    const bool last_level = (e == p.emax);
    CarryT br[ITEMS];
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
        const int t = t0 + r * THREADS + tid;
        const int seg = kb + s_wpre[r * WARPS + warp] + __popc(bal[r] & (0xffffffffu >> (31 - lane)));
        const int jj = min(seg, K) - lo;
        CarryT b = A::add(kL[jj], A::mul(kS[jj], A::sub(xr[r], kX[jj])));
        if (t >= n - 1) b = (CarryT)0;                  // ITD.py:112: sample N-1 is never written
        br[r] = b;
        if (t < n) {
            const CarryT rr = A::sub(xr[r], b);
            // iteration stop: the last row is rotation + baseline (ITD.py:420)
            rot[t] = (OutT)(last_level ? A::add(rr, b) : rr);
            carry[t] = b;
            // on the iteration stop the reference's last baseline row is never written (ITD.py:424)
            if (bas) bas[t] = last_level ? (OutT)0 : (OutT)b;
        }
    }
    // halo samples of B, needed by the stencil on B at the tile edges
    CarryT bl = (CarryT)0, bq = (CarryT)0;
    if (tid == 0 && t0 > 0) {
        const int jj = min(kb, K) - lo;
        bl = A::add(kL[jj], A::mul(kS[jj], A::sub(xs[0], kX[jj])));
    }
    if (tid == 32 % THREADS && t0 + T < n - 1) {
        const int t = t0 + T;                           // first sample of the next tile
        const bool f = (t <= n - 2) && is_knot(xs[T], xs[T + 1], xs[T + 2]);
        const int jj = min(kb + cnt + (f ? 1 : 0), K) - lo;
        bq = A::add(kL[jj], A::mul(kS[jj], A::sub(xs[T + 1], kX[jj])));
    }
    __syncthreads();        // every read of xs / kL / kS is done
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) xs[1 + r * THREADS + tid] = br[r];
    if (tid == 0) xs[0] = bl;
    if (tid == 32 % THREADS) xs[T + 1] = bq;
    __syncthreads();

    // ---- 6. extrema of B = the stop test (ITD.py:400-404) = the next level's knots -------------
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
        const int j = 1 + r * THREADS + tid, t = t0 + r * THREADS + tid;
        const bool f = (t >= 1) && (t <= n - 2) && is_knot(xs[j - 1], xs[j], xs[j + 1]);
        bal[r] = __ballot_sync(0xffffffffu, f);
    }
    int total;
    const int base = tile_scan<THREADS, ITEMS>(bal, s_wcnt, s_wpre, s_misc,
                                               p.desc + (long long)sig * p.tiles, tile, p.tag, total);
    int *tau = p.next.tau + (long long)sig * p.next.kstride;
    CarryT *xk = reinterpret_cast<CarryT *>(p.next.xk) + (long long)sig * p.next.kstride;
    unsigned *gmask = p.next.mask + (long long)sig * p.next.mstride + tile * (T / 32);
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
        if (lane == 0 && t0 + r * THREADS + warp * 32 < n) gmask[r * WARPS + warp] = bal[r];
        if ((bal[r] >> lane) & 1u) {
            const int rank = base + s_wpre[r * WARPS + warp] + __popc(bal[r] & ((1u << lane) - 1u));
            tau[1 + rank] = t0 + r * THREADS + tid;
            xk[1 + rank] = br[r];
        }
    }
    if (tid == 0) {
        int *tb = p.next.tbase + (long long)sig * (p.tiles + 1);
        CarryT *endl = reinterpret_cast<CarryT *>(p.next.endl) + 2ll * sig;
        tb[tile] = base;
        if (tile == 0) {
            tau[0] = 0;
            xk[0] = xs[1];
            endl[0] = mean2<CarryT>(xs[1], xs[2]);
        }
        if (tile == p.tiles - 1) {
            const int Kn = base + total;
            const int jl = n - 1 - t0 + 1;
            tb[p.tiles] = Kn;
            p.next.kcount[sig] = Kn;
            tau[Kn + 1] = n - 1;
            xk[Kn + 1] = xs[jl];
            endl[1] = mean2<CarryT>(xs[jl - 1], xs[jl]);
            p.knot_counts[(long long)sig * p.rows + e] = Kn;      // what ITD.py:403 prints
            if (Kn < p.min_extrema) {                             // ITD.py:404
                p.stop_kind[sig] = kStopKnots;
                p.n_rows[sig] = e + 1;
                p.stop_e[sig] = e;
            } else if (last_level) {                              // ITD.py:418
                p.stop_kind[sig] = kStopIter;
                p.n_rows[sig] = e + 1;
                p.stop_e[sig] = e;
            }
        }
    }
}

// copies the interior knots of the current table into a user buffer (find_knots entry point)
__global__ void export_knots_kernel(const int *tau, const int *kcount, long long kstride, int *out,
                                    long long capacity, int *count_out) {
    const int sig = blockIdx.y;
    const int K = kcount[sig];
    if (blockIdx.x == 0 && threadIdx.x == 0) count_out[sig] = K;
    const long long lim = (K < capacity) ? K : capacity;
    const int *src = tau + (long long)sig * kstride + 1;
    int *dst = out + (long long)sig * capacity;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < lim;
         i += (long long)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

// ---------------------------------------------------------------------------------------------
// table_from_knots_kernel: the knot table of a level from a SUPPLIED knot list instead of a scan of x
// (the reference's C++ port keeps and reuses the extrema "along multiple channels", itd.cpp:41-44, with
// compute_extrema == false at itd.cpp:156-169; here on ITD.py's own interpolant, ITD.py:95-119).
// One CTA per signal.  knots: [rows, cap] ascending interior indices, rows = S or 1 (shared list).
// Writes exactly what knot_scan_kernel writes: tau / X_k (with the end knots 0 and n-1, ITD.py:98), the
// flag mask, the per-tile knot prefix, K and the two end baselines (ITD.py:101-102).  A list that is not
// strictly increasing inside [1, n-2] is replaced by the empty list and reported as kStBadKnots.
// ---------------------------------------------------------------------------------------------
template <typename InT, typename CarryT>
__global__ void __launch_bounds__(256) table_from_knots_kernel(const void *xin, const int *knots, long long cap,
                                                               const int *counts, int shared_list, KnotTable out,
                                                               int *status, int n, int tiles, int tile) {
    const int sig = blockIdx.x, tid = threadIdx.x;
    const long long kr = shared_list ? 0 : sig;
    const InT *x = reinterpret_cast<const InT *>(xin) + (long long)sig * n;
    const int *kn = knots + kr * cap;
    int K = counts[kr];
    bool bad = (K < 0) || ((long long)K > cap) || (K > n - 2);
    if (bad) K = 0;
    bool nonfinite = false;
    for (int t = tid; t < n; t += blockDim.x) nonfinite |= !isfinite((CarryT)x[t]);
    for (int j = tid; j < K; j += blockDim.x) {
        const int t = kn[j];
        bad |= (t < 1) || (t > n - 2) || (j > 0 && t <= kn[j - 1]);
    }
    bad = __syncthreads_or(bad ? 1 : 0) != 0;
    nonfinite = __syncthreads_or(nonfinite ? 1 : 0) != 0;
    if (bad) K = 0;
    if (tid == 0 && (bad || nonfinite)) atomicOr(status + sig, (bad ? kStBadKnots : 0) | (nonfinite ? kStNonFinite : 0));

    int *tau = out.tau + (long long)sig * out.kstride;
    CarryT *xk = reinterpret_cast<CarryT *>(out.xk) + (long long)sig * out.kstride;
    unsigned *mask = out.mask + (long long)sig * out.mstride;
    int *tbase = out.tbase + (long long)sig * (tiles + 1);
    for (long long w = tid; w < out.mstride; w += blockDim.x) mask[w] = 0u;
    for (int j = tid; j < K; j += blockDim.x) {
        const int t = kn[j];
        tau[1 + j] = t;
        xk[1 + j] = (CarryT)x[t];
    }
    if (tid == 0) {
        const CarryT a0 = (CarryT)x[0], a1 = (CarryT)x[1], z1 = (CarryT)x[n - 2], z0 = (CarryT)x[n - 1];
        tau[0] = 0;
        xk[0] = a0;
        tau[K + 1] = n - 1;
        xk[K + 1] = z0;
        out.kcount[sig] = K;
        CarryT *endl = reinterpret_cast<CarryT *>(out.endl) + 2ll * sig;
        endl[0] = mean2<CarryT>(a0, a1);
        endl[1] = mean2<CarryT>(z1, z0);
    }
    // knots before each tile: lower bound of the tile's first sample in the list
    for (int i = tid; i <= tiles; i += blockDim.x) {
        const long long target = (long long)i * tile;
        int lo = 0, hi = K;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (kn[mid] < target) lo = mid + 1;
            else hi = mid;
        }
        tbase[i] = (i == tiles) ? K : lo;
    }
    __syncthreads();                                   // the mask row is cleared
    for (int j = tid; j < K; j += blockDim.x) {
        const int t = kn[j];
        atomicOr(mask + (t >> 5), 1u << (t & 31));
    }
}

template <int THREADS, int ITEMS, typename CarryT>
constexpr size_t level_smem_bytes() {
    return sizeof(CarryT) * (size_t)(THREADS * ITEMS + 4) + 3 * sizeof(CarryT) * (size_t)(THREADS * ITEMS + 8) +
           sizeof(int) * (size_t)(THREADS * ITEMS + 8) + sizeof(int) * (size_t)(2 * (THREADS * ITEMS / 32) + 8);
}
template <int THREADS, int ITEMS, typename CarryT>
constexpr size_t scan_smem_bytes() {
    return sizeof(CarryT) * (size_t)(THREADS * ITEMS + 4) + sizeof(int) * (size_t)(2 * (THREADS * ITEMS / 32) + 8);
}

}  // namespace pyitd
