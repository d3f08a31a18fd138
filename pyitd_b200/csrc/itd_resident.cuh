// itd_resident.cuh -- the whole decomposition of a signal in ONE kernel, with the signal resident ON CHIP.
//
// One thread-block cluster (1..8 CTAs) owns one signal at a time.  Every CTA keeps its chunk of the
// carry X_e in shared memory for all levels, so per level the only HBM traffic is the rotation row
// going out (plus the optional baseline row): read x once, write each output row once.  The level
// loop, the stop test (ITD.py:400-404, :418) and the trend-row fix-up (ITD.py:410-411) all run inside
// the kernel -- there is one launch per batch and no host synchronisation.
//
// Work decomposition
//   * a UNIT is 32*SPL consecutive samples: lane l of a warp owns samples [l*SPL, (l+1)*SPL) of it
//     (blocked layout: the 3-point stencil sees its neighbours in registers, loads and stores are
//     128-bit, the carry array is XOR-swizzled per 16-byte chunk so those accesses are conflict free);
//   * the units of a signal are dealt out contiguously to the CL*WARPS warps of the cluster; a warp
//     walks its range left to right and nothing but that warp ever touches that part of the carry or
//     of the knot-flag masks -- no block barrier inside a level;
//   * knots are kept as a bit mask (one bit per sample).  For a RUN of units holding at most UNIT knots
//     the warp enumerates the knots into a private table, adds the two knots before and three after
//     (from its own mask ahead or from the neighbours' published summaries), and evaluates the knot
//     baseline L_k (ITD.py:100-110) and the segment slopes (ITD.py:116) with one lane per knot;
//   * then per sample: B = L_k + s_k (x - X_k), R = x - B (ITD.py:115-119), R -> HBM, B -> carry, and
//     the extrema of B (= the stop test = the next level's knots) as a new bit mask;
//   * per level every warp publishes {knot count, first three knots, last two knots} of its range;
//     a CTA aggregates its warps' summaries for the other CTAs to read over DSMEM; one cluster
//     barrier per level.
//
// X_e must survive until the stop test on B_e is known (the knot stop returns X_e as the trend row,
// ITD.py:410-411) although the carry is updated in place: each level first saves its input to a
// per-cluster backup area (L2-resident scratch in global memory).
//
// fp64 arithmetic uses the unfused intrinsics of itd_kernels.cuh in the reference's operation order.
#pragma once

#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "itd_kernels.cuh"

namespace pyitd {

namespace cg = cooperative_groups;

struct ResidentParams {
    const void *x;              // [S, n] input type
    void *rot;                  // [S, rows, n] output type
    void *bas;                  // [S, rows, n] output type or null
    void *backup;               // [clusters, backup_stride] carry type
    long long out_sig_stride;   // rows * n
    long long backup_stride;    // nu * UNIT
    int *n_rows, *knot_counts, *input_knots, *stop_kind, *status;
    long long S;
    int n;
    int nu;                     // units per signal = ceil(n / UNIT)
    int chunk_units;            // shared-memory capacity of a CTA, in units
    int emax, rows, min_extrema;
    unsigned opts;
    // shared memory byte offsets
    unsigned off_mask, off_tab, off_ws, off_cs, off_misc;
};

// first unit of global warp g when nu units are dealt to gw warps
__host__ __device__ inline int res_unit_begin(int g, int gw, int nu) {
    return (int)(((long long)g * (long long)nu) / (long long)gw);
}

template <typename CarryT>
struct alignas(16) KnotSummary {
    int cnt;          // knots in the range
    int tF[3];        // first three knot positions (ascending)
    int tL[2];        // last knot, second-last knot
    int pad_[2];
    CarryT xF[3];     // signal values at those knots
    CarryT xL[2];
    CarryT padv_[1];
};
template <typename CarryT>
struct alignas(16) LevelMisc {
    CarryT endl0, endl1;   // L_0 and L_{K+1} (ITD.py:101-102)
    CarryT x0, xlast;      // x[0] and x[n-1] of this level (the two virtual end knots' values)
};
template <typename CarryT>
struct alignas(2 * sizeof(CarryT)) KnotLS {
    CarryT L, s;
};

// the resolved neighbourhood of a warp's range: two knots before (nearest first), three after
template <typename CarryT>
struct alignas(16) HaloKnots {
    int bt[2], at[3], pad_[3];
    CarryT bx[2], ax[3], padv_[1];
};

template <typename CarryT, int WARPS, int SPL>
struct ResidentGeom {
    static constexpr int UNIT = 32 * SPL;
    static constexpr int CAP = UNIT + 8;                       // knot-table capacity per warp
    static constexpr size_t tab_core_bytes =
        ((size_t)CAP * (sizeof(int) + sizeof(CarryT) + sizeof(KnotLS<CarryT>)) + 15) & ~(size_t)15;
    static constexpr size_t tab_bytes_per_warp = tab_core_bytes + sizeof(HaloKnots<CarryT>);
    // byte offsets for a CTA holding chunk_units units
    static void layout(int chunk_units, ResidentParams &p, size_t &total) {
        size_t o = (size_t)chunk_units * UNIT * sizeof(CarryT);
        o = (o + 127) & ~(size_t)127;
        p.off_mask = (unsigned)o;
        o += 2ull * chunk_units * SPL * sizeof(unsigned);
        o = (o + 15) & ~(size_t)15;
        p.off_tab = (unsigned)o;
        o += tab_bytes_per_warp * WARPS;
        p.off_ws = (unsigned)o;
        o += 2ull * WARPS * sizeof(KnotSummary<CarryT>);
        p.off_cs = (unsigned)o;
        o += 2ull * sizeof(KnotSummary<CarryT>);
        p.off_misc = (unsigned)o;
        o += 2ull * sizeof(LevelMisc<CarryT>);
        total = o;
    }
};

// ---------------------------------------------------------------------------------------------
// comparison bits.  cmp_bits ORs `mask` into lt when a < b and into gt when a > b (IEEE compare:
// -0 == +0, exactly what numpy's `>` / `<=` on the differences do in ITD.py:59).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cmp_bits(double a, double b, unsigned mask, unsigned &lt, unsigned &gt) {
    asm("{\n\t.reg .pred p, q;\n\t"
        "setp.lt.f64 p, %2, %3;\n\t"
        "setp.gt.f64 q, %2, %3;\n\t"
        "@p or.b32 %0, %0, %4;\n\t"
        "@q or.b32 %1, %1, %4;\n\t}"
        : "+r"(lt), "+r"(gt)
        : "d"(a), "d"(b), "r"(mask));
}
__device__ __forceinline__ void cmp_bits(float a, float b, unsigned mask, unsigned &lt, unsigned &gt) {
    asm("{\n\t.reg .pred p, q;\n\t"
        "setp.lt.f32 p, %2, %3;\n\t"
        "setp.gt.f32 q, %2, %3;\n\t"
        "@p or.b32 %0, %0, %4;\n\t"
        "@q or.b32 %1, %1, %4;\n\t}"
        : "+r"(lt), "+r"(gt)
        : "f"(a), "f"(b), "r"(mask));
}

// per-lane knot flags of SPL consecutive values.  lt/gt bit j = comparison of sample j-1 with sample j
// (j = 0 compares the left neighbour vl).  The flag of the lane's last sample needs the next lane's
// first comparison (one shuffle); lane 31's last flag is left to the caller (deferred to the next unit).
template <int SPL, typename CarryT>
__device__ __forceinline__ unsigned lane_flags(const CarryT (&v)[SPL], CarryT vl, int lane, unsigned &pk_first,
                                               unsigned &pk_last) {
    unsigned lt = 0u, gt = 0u;
    cmp_bits(vl, v[0], 1u, lt, gt);
#pragma unroll
    for (int j = 1; j < SPL; ++j) cmp_bits(v[j - 1], v[j], 1u << j, lt, gt);
    pk_first = (lt & 1u) | ((gt & 1u) << 1);
    pk_last = ((lt >> (SPL - 1)) & 1u) | (((gt >> (SPL - 1)) & 1u) << 1);
    const unsigned nx = __shfl_down_sync(0xffffffffu, pk_first, 1);
    const unsigned lte = lt | ((nx & 1u) << SPL), gte = gt | ((nx >> 1) << SPL);
    // valley: !(x[t-1] < x[t]) && x[t] < x[t+1];  peak: !(x[t-1] > x[t]) && x[t] > x[t+1]   (ITD.py:59 on x and -x)
    unsigned f = ((~lt) & (lte >> 1)) | ((~gt) & (gte >> 1));
    f &= (lane == 31) ? ((1u << (SPL - 1)) - 1u) : ((1u << SPL) - 1u);
    return f;
}
__device__ __forceinline__ unsigned flag_from_pk(unsigned pk_left, unsigned pk_right) {
    // pk = (lt | gt << 1) of (t-1, t) resp. (t, t+1)
    const unsigned m = ~pk_left & pk_right;
    return (m | (m >> 1)) & 1u;
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <typename InT, typename CarryT, typename OutT, int WARPS, int SPL>
__global__ void __launch_bounds__(WARPS * 32, 1) resident_kernel(const ResidentParams p) {
    using A = Arith<CarryT>;
    using G = ResidentGeom<CarryT, WARPS, SPL>;
    using Summary = KnotSummary<CarryT>;
    using Misc = LevelMisc<CarryT>;
    using LS = KnotLS<CarryT>;
    using Halo = HaloKnots<CarryT>;
    using CVec = typename std::conditional<sizeof(CarryT) == 8, double2, float4>::type;
    using OVec = typename std::conditional<sizeof(OutT) == 8, double2, float4>::type;
    constexpr int UNIT = G::UNIT, CAP = G::CAP;
    constexpr int EPC = 16 / (int)sizeof(CarryT);          // carry elements per 16-byte chunk
    constexpr int LCH = SPL / EPC;                         // chunks per lane per unit
    constexpr int OPV = 16 / (int)sizeof(OutT);            // output elements per 16-byte store
    constexpr int LPW = 32 / SPL;                          // lanes per mask word
    constexpr unsigned FBM = (1u << SPL) - 1u;
    static_assert(SPL == 4 || SPL == 8, "SPL must be 4 or 8");
    static_assert(LCH >= 1 && LCH <= 4, "a lane owns 1..4 16-byte chunks of the carry per unit");
    constexpr unsigned FULL = 0xffffffffu;

    extern __shared__ __align__(128) unsigned char smem_res[];
    cg::cluster_group cluster = cg::this_cluster();
    const int CL = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int cid = (int)(blockIdx.x / CL), ncl = (int)(gridDim.x / CL);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int GW = CL * WARPS, g = rank * WARPS + warp;
    const int n = p.n, nu = p.nu;

    // ---- geometry of this warp's range ---------------------------------------------------------
    const int u0 = res_unit_begin(g, GW, nu), u1 = res_unit_begin(g + 1, GW, nu);
    const int cu0 = res_unit_begin(rank * WARPS, GW, nu);                 // first unit of the CTA's chunk
    const bool have = u0 < u1;
    const int a = u0 * UNIT, b = have ? min(u1 * UNIT, n) : a;            // samples [a, b)

    CarryT *Xs = reinterpret_cast<CarryT *>(smem_res);
    unsigned *Mk = reinterpret_cast<unsigned *>(smem_res + p.off_mask);   // [2][chunk_units * SPL]
    unsigned char *tab = smem_res + p.off_tab + (size_t)warp * G::tab_bytes_per_warp;
    LS *ls = reinterpret_cast<LS *>(tab);
    CarryT *XT = reinterpret_cast<CarryT *>(tab + (size_t)CAP * sizeof(LS));
    int *TAU = reinterpret_cast<int *>(tab + (size_t)CAP * (sizeof(LS) + sizeof(CarryT)));
    Halo *hk = reinterpret_cast<Halo *>(tab + G::tab_core_bytes);
    Summary *WS = reinterpret_cast<Summary *>(smem_res + p.off_ws);       // [2][WARPS]
    Summary *CS = reinterpret_cast<Summary *>(smem_res + p.off_cs);       // [2]
    Misc *MISC = reinterpret_cast<Misc *>(smem_res + p.off_misc);         // [2], the copy in CTA 0 is the live one
    Misc *MISC0 = cluster.map_shared_rank(MISC, 0);
    const int mwords = p.chunk_units * SPL;

    // per-lane constants
    const int lw = lane / LPW, lsh = (lane % LPW) * SPL;
    const int xorv = ((lane * LCH) >> 3) & (LCH - 1);
    auto swz = [&](int trel) -> int {                                     // element index in Xs of chunk-relative sample
        const int c = trel / EPC;
        return ((c ^ ((c >> 3) & (LCH - 1))) * EPC) + (trel % EPC);
    };
    // element offsets (inside a unit) of the lane's LCH chunks
    int coff[LCH];
#pragma unroll
    for (int qc = 0; qc < LCH; ++qc) coff[qc] = (lane * LCH + (qc ^ xorv)) * EPC;

    int par = 0;                       // parity of the mask / summary buffers holding the CURRENT level's knots
    CarryT *bk = reinterpret_cast<CarryT *>(p.backup) + (long long)cid * p.backup_stride;

    for (long long sig = cid; sig < p.S; sig += ncl) {
        const InT *x = reinterpret_cast<const InT *>(p.x) + sig * n;
        OutT *rot = reinterpret_cast<OutT *>(p.rot) + sig * p.out_sig_stride;
        OutT *bas = p.bas ? reinterpret_cast<OutT *>(p.bas) + sig * p.out_sig_stride : nullptr;
        const bool in_vec = ((reinterpret_cast<uintptr_t>(x) | ((size_t)n * sizeof(InT))) & 15) == 0 && (SPL * sizeof(InT)) % 16 == 0;
        const bool out_vec = ((reinterpret_cast<uintptr_t>(rot) | ((size_t)n * sizeof(OutT))) & 15) == 0 &&
                             (!bas || (reinterpret_cast<uintptr_t>(bas) & 15) == 0) && (SPL * sizeof(OutT)) % 16 == 0;
        bool bad = false, zero_dx = false;
        par ^= 1;        // slow warps may still be reading the previous signal's last summaries (buffer par)
        CarryT hxl = (CarryT)0, hxr = (CarryT)0;      // values of samples a-1 and b at the current level

        // =====================================================================================
        // helpers
        // =====================================================================================
        // flag bits of the lane -> mask word of unit `cu` (chunk-relative) in buffer mkq
        auto store_flags = [&](unsigned *mkq, int cu, unsigned f) {
            unsigned w = f << lsh;
#pragma unroll
            for (int o = 1; o < LPW; o <<= 1) w |= __shfl_xor_sync(FULL, w, o);
            if ((lane % LPW) == 0) mkq[cu * SPL + lw] = w;
        };
        // first three / last two set bits of my range in mask buffer mkq (global sample positions, -1 = none)
        auto range_ends = [&](const unsigned *mkq, int &tF0, int &tF1, int &tF2, int &tL0, int &tL1) {
            const int w0 = (u0 - cu0) * SPL, w1 = (u1 - cu0) * SPL;
            int found = 0;
            for (int wb = w0; wb < w1 && found < 3; wb += 32) {
                const unsigned wd = (wb + lane < w1) ? mkq[wb + lane] : 0u;
                unsigned nz = __ballot_sync(FULL, wd != 0u);
                while (nz && found < 3) {
                    const int fl = __ffs(nz) - 1;
                    nz &= nz - 1;
                    unsigned ww = __shfl_sync(FULL, wd, fl);
                    while (ww && found < 3) {
                        const int t = (cu0 * SPL + wb + fl) * 32 + (__ffs(ww) - 1);
                        ww &= ww - 1;
                        if (found == 0) tF0 = t; else if (found == 1) tF1 = t; else tF2 = t;
                        ++found;
                    }
                }
            }
            found = 0;
            for (int we = w1; we > w0 && found < 2; we -= 32) {
                const int wi = we - 1 - lane;                       // lane 0 = highest word
                const unsigned wd = (wi >= w0) ? mkq[wi] : 0u;
                unsigned nz = __ballot_sync(FULL, wd != 0u);
                while (nz && found < 2) {
                    const int fl = __ffs(nz) - 1;
                    nz &= nz - 1;
                    unsigned ww = __shfl_sync(FULL, wd, fl);
                    while (ww && found < 2) {
                        const int hb = 31 - __clz(ww);
                        const int t = (cu0 * SPL + we - 1 - fl) * 32 + hb;
                        ww &= ~(1u << hb);
                        if (found == 0) tL0 = t; else tL1 = t;
                        ++found;
                    }
                }
            }
        };
        // publish this warp's summary of mask buffer q: count + first three / last two knots with values
        auto publish = [&](int q, int cnt_total) {
            Summary *me = &WS[q * WARPS + warp];
            int tF0 = -1, tF1 = -1, tF2 = -1, tL0 = -1, tL1 = -1;
            if (cnt_total > 0) range_ends(Mk + q * mwords, tF0, tF1, tF2, tL0, tL1);
            if (lane == 0) {
                me->cnt = cnt_total;
                if (cnt_total > 0) {
                    me->tF[0] = tF0; me->tF[1] = tF1; me->tF[2] = tF2;
                    me->tL[0] = tL0; me->tL[1] = tL1;
                    me->xF[0] = Xs[swz(tF0 - cu0 * UNIT)];
                    me->xL[0] = Xs[swz(tL0 - cu0 * UNIT)];
                    if (tF1 >= 0) me->xF[1] = Xs[swz(tF1 - cu0 * UNIT)];
                    if (tF2 >= 0) me->xF[2] = Xs[swz(tF2 - cu0 * UNIT)];
                    if (tL1 >= 0) me->xL[1] = Xs[swz(tL1 - cu0 * UNIT)];
                }
            }
        };
        // block barrier, CTA aggregate for the other CTAs, cluster barrier
        auto level_sync = [&](int q) {
            if (CL > 1) {
                __syncthreads();
                if (warp == 0) {
                    const Summary *src = &WS[q * WARPS];
                    const int c = (lane < WARPS) ? src[lane].cnt : 0;
                    const int tot = __reduce_add_sync(FULL, c);
                    int tF0 = -1, tF1 = -1, tF2 = -1, tL0 = -1, tL1 = -1;
                    CarryT xF0 = 0, xF1 = 0, xF2 = 0, xL0 = 0, xL1 = 0;
                    const unsigned m = __ballot_sync(FULL, c > 0);
                    int found = 0;
                    unsigned mm = m;
                    while (mm && found < 3) {
                        const int j = __ffs(mm) - 1;
                        mm &= mm - 1;
                        const int cj = min(src[j].cnt, 3);
                        for (int i = 0; i < cj && found < 3; ++i, ++found) {
                            const int t = src[j].tF[i];
                            const CarryT v = src[j].xF[i];
                            if (found == 0) { tF0 = t; xF0 = v; } else if (found == 1) { tF1 = t; xF1 = v; } else { tF2 = t; xF2 = v; }
                        }
                    }
                    found = 0;
                    mm = m;
                    while (mm && found < 2) {
                        const int j = 31 - __clz(mm);
                        mm &= ~(1u << j);
                        const int cj = min(src[j].cnt, 2);
                        for (int i = 0; i < cj && found < 2; ++i, ++found) {
                            const int t = src[j].tL[i];
                            const CarryT v = src[j].xL[i];
                            if (found == 0) { tL0 = t; xL0 = v; } else { tL1 = t; xL1 = v; }
                        }
                    }
                    if (lane == 0) {
                        Summary *d = &CS[q];
                        d->cnt = tot;
                        d->tF[0] = tF0; d->tF[1] = tF1; d->tF[2] = tF2; d->tL[0] = tL0; d->tL[1] = tL1;
                        d->xF[0] = xF0; d->xF[1] = xF1; d->xF[2] = xF2; d->xL[0] = xL0; d->xL[1] = xL1;
                    }
                }
            }
            cluster.sync();
        };

        // neighbourhood of this warp's range at the current level: scalars in registers, knots in hk
        int K = 0, kb = 0, mycnt = 0;
        CarryT endl0 = 0, endl1 = 0;
        auto resolve = [&](int q) {
            // entry list in sample order: CTAs before mine (aggregates), my CTA's warps, CTAs after mine
            const int ne = CL - 1 + WARPS, me = rank + warp;
            const Summary *e = nullptr;
            if (lane < rank) e = cluster.map_shared_rank(&CS[q], lane);
            else if (lane < rank + WARPS) e = &WS[q * WARPS + (lane - rank)];
            else if (lane < ne) e = cluster.map_shared_rank(&CS[q], lane - WARPS + 1);
            const int c = e ? e->cnt : 0;
            int tF0 = 0, tF1 = 0, tF2 = 0, tL0 = 0, tL1 = 0;
            CarryT xF0 = 0, xF1 = 0, xF2 = 0, xL0 = 0, xL1 = 0;
            if (c > 0) {
                tF0 = e->tF[0]; tF1 = e->tF[1]; tF2 = e->tF[2]; tL0 = e->tL[0]; tL1 = e->tL[1];
                xF0 = e->xF[0]; xF1 = e->xF[1]; xF2 = e->xF[2]; xL0 = e->xL[0]; xL1 = e->xL[1];
            }
            const Misc mi = MISC0[q];
            endl0 = mi.endl0; endl1 = mi.endl1;
            const CarryT x0v = mi.x0, xlastv = mi.xlast;
            K = __reduce_add_sync(FULL, c);
            kb = __reduce_add_sync(FULL, (lane < me) ? c : 0);
            mycnt = __shfl_sync(FULL, c, me);
            const unsigned nzm = __ballot_sync(FULL, c > 0);
            // two nearest real knots before the range
            int nb = 0;
            unsigned m = nzm & ((1u << me) - 1u);
            int rt0 = 0, rt1 = 0;
            CarryT rx0 = 0, rx1 = 0;
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                if (m && nb < 2) {
                    const int j = 31 - __clz(m);
                    m &= ~(1u << j);
                    const int cj = __shfl_sync(FULL, c, j);
                    const int t0_ = __shfl_sync(FULL, tL0, j), t1_ = __shfl_sync(FULL, tL1, j);
                    const CarryT v0_ = __shfl_sync(FULL, xL0, j), v1_ = __shfl_sync(FULL, xL1, j);
                    if (nb == 0) { rt0 = t0_; rx0 = v0_; } else { rt1 = t0_; rx1 = v0_; }
                    ++nb;
                    if (nb < 2 && cj >= 2) { rt1 = t1_; rx1 = v1_; ++nb; }
                }
            }
            // three nearest real knots after the range
            int na = 0;
            m = (me >= 31) ? 0u : (nzm & ~((2u << me) - 1u));
            int qt0 = 0, qt1 = 0, qt2 = 0;
            CarryT qx0 = 0, qx1 = 0, qx2 = 0;
#pragma unroll
            for (int it = 0; it < 3; ++it) {
                if (m && na < 3) {
                    const int j = __ffs(m) - 1;
                    m &= m - 1;
                    const int cj = __shfl_sync(FULL, c, j);
                    const int t0_ = __shfl_sync(FULL, tF0, j), t1_ = __shfl_sync(FULL, tF1, j), t2_ = __shfl_sync(FULL, tF2, j);
                    const CarryT v0_ = __shfl_sync(FULL, xF0, j), v1_ = __shfl_sync(FULL, xF1, j), v2_ = __shfl_sync(FULL, xF2, j);
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        if (i < cj && na < 3) {
                            const int t = (i == 0) ? t0_ : (i == 1) ? t1_ : t2_;
                            const CarryT v = (i == 0) ? v0_ : (i == 1) ? v1_ : v2_;
                            if (na == 0) { qt0 = t; qx0 = v; } else if (na == 1) { qt1 = t; qx1 = v; } else { qt2 = t; qx2 = v; }
                            ++na;
                        }
                    }
                }
            }
            if (lane == 0) {
                // knot kb (nearest before) and kb-1: real, or the virtual start knot (tau 0, x[0]), or unused
                hk->bt[0] = (kb >= 1) ? rt0 : 0; hk->bx[0] = (kb >= 1) ? rx0 : x0v;
                hk->bt[1] = (kb >= 2) ? rt1 : 0; hk->bx[1] = (kb >= 2) ? rx1 : x0v;
                // knots ka+1.. (ka = kb + mycnt): real while <= K, then the virtual end knot (tau n-1, x[n-1])
                const int ka = kb + mycnt;
                hk->at[0] = (ka + 1 <= K) ? qt0 : n - 1; hk->ax[0] = (ka + 1 <= K) ? qx0 : xlastv;
                hk->at[1] = (ka + 2 <= K) ? qt1 : n - 1; hk->ax[1] = (ka + 2 <= K) ? qx1 : xlastv;
                hk->at[2] = (ka + 3 <= K) ? qt2 : n - 1; hk->ax[2] = (ka + 3 <= K) ? qx2 : xlastv;
            }
            __syncwarp();
        };

        // per-warp streaming state of a pass over the range
        int cnt_lane = 0;
        unsigned pend = 0;             // pk_last of lane 31 of the previous unit
        CarryT vlast = (CarryT)0;      // value of the previous unit's last sample
        // flags of one unit from its values (in registers) -> mask buffer mkq; counts into cnt_lane
        auto emit_flags = [&](auto edge_c, const CarryT (&v)[SPL], unsigned *mkq, int uu) {
            constexpr bool EDGE = decltype(edge_c)::value;
            const int cu = uu - cu0, t0 = uu * UNIT;
            CarryT vl = __shfl_up_sync(FULL, v[SPL - 1], 1);
            if (lane == 0) vl = vlast;
            unsigned pkf, pkl;
            unsigned f = lane_flags<SPL, CarryT>(v, vl, lane, pkf, pkl);
            if (EDGE) {
                const int tl = t0 + lane * SPL;
                const int lo = max(0, 1 - tl), hi = n - 2 - tl;            // valid bits [lo, hi]
                unsigned vm = (hi < 0) ? 0u : ((hi >= 31) ? FULL : ((2u << hi) - 1u));
                vm &= (lo >= 32) ? 0u : (FULL << lo);
                f &= vm;
            }
            // deferred flag of the previous unit's last sample (position t0 - 1, inside this range)
            const unsigned pk0 = __shfl_sync(FULL, pkf, 0);
            if (uu > u0 && flag_from_pk(pend, pk0) && (!EDGE || (t0 - 1 >= 1 && t0 - 1 <= n - 2))) {
                if (lane == 0) {
                    mkq[(cu - 1) * SPL + SPL - 1] |= 0x80000000u;
                    ++cnt_lane;
                }
            }
            pend = __shfl_sync(FULL, pkl, 31);
            vlast = __shfl_sync(FULL, v[SPL - 1], 31);
            cnt_lane += __popc(f);
            store_flags(mkq, cu, f);
            __syncwarp();
        };
        // the range's very last sample: its right neighbour (value hr) lives in the next warp's range
        auto close_range = [&](unsigned *mkq, CarryT hr) {
            if (have && b < n && lane == 0) {
                const unsigned pkr = (vlast < hr ? 1u : 0u) | (vlast > hr ? 2u : 0u);
                if (b - 1 >= 1 && b - 1 <= n - 2 && flag_from_pk(pend, pkr)) {
                    mkq[(u1 - 1 - cu0) * SPL + SPL - 1] |= 0x80000000u;
                    ++cnt_lane;
                }
            }
            __syncwarp();
        };

        // =====================================================================================
        // 0. load the chunk, detect the extrema of the input (ITD.py:87-98) -> mask[par]
        // =====================================================================================
        {
            unsigned *mkq = Mk + par * mwords;
            if (have) {
                if (a > 0) hxl = (CarryT)__ldg(x + a - 1);
                if (b < n) hxr = (CarryT)__ldg(x + b);
                vlast = hxl;
            }
            for (int u = u0; u < u1; ++u) {
                const int t0 = u * UNIT, tl = t0 + lane * SPL, cu = u - cu0;
                const bool edge = (u == 0) || (t0 + UNIT >= n - 1);       // holds sample 0, n-2 or n-1
                CarryT v[SPL];
                if (!edge && in_vec) {
                    constexpr int IPV = 16 / (int)sizeof(InT);
#pragma unroll
                    for (int qv = 0; qv < SPL / IPV; ++qv) {
                        if constexpr (sizeof(InT) == 8) {
                            const double2 d = __ldg(reinterpret_cast<const double2 *>(x + tl) + qv);
                            v[qv * 2] = (CarryT)d.x; v[qv * 2 + 1] = (CarryT)d.y;
                        } else {
                            const float4 d = __ldg(reinterpret_cast<const float4 *>(x + tl) + qv);
                            v[qv * 4] = (CarryT)d.x; v[qv * 4 + 1] = (CarryT)d.y; v[qv * 4 + 2] = (CarryT)d.z; v[qv * 4 + 3] = (CarryT)d.w;
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < SPL; ++j) v[j] = (tl + j < n) ? (CarryT)__ldg(x + tl + j) : (CarryT)0;
                }
#pragma unroll
                for (int j = 0; j < SPL; ++j) bad |= !isfinite(v[j]);
                // carry <- x (swizzled 16-byte chunks)
                CarryT *xu = Xs + (size_t)cu * UNIT;
#pragma unroll
                for (int qc = 0; qc < LCH; ++qc) {
                    CVec o;
                    if constexpr (sizeof(CarryT) == 8) { o.x = v[qc * 2]; o.y = v[qc * 2 + 1]; }
                    else { o.x = v[qc * 4]; o.y = v[qc * 4 + 1]; o.z = v[qc * 4 + 2]; o.w = v[qc * 4 + 3]; }
                    *reinterpret_cast<CVec *>(xu + coff[qc]) = o;
                }
                if (edge) emit_flags(std::true_type{}, v, mkq, u);
                else emit_flags(std::false_type{}, v, mkq, u);
            }
            close_range(mkq, hxr);
            const int cnt_total = __reduce_add_sync(FULL, cnt_lane);
            publish(par, cnt_total);
            if (g == 0 && lane == 0) {
                Misc mi;
                const CarryT xa = (CarryT)__ldg(x), xb = (CarryT)__ldg(x + 1);
                const CarryT xy = (CarryT)__ldg(x + n - 2), xz = (CarryT)__ldg(x + n - 1);
                mi.endl0 = mean2<CarryT>(xa, xb);
                mi.endl1 = mean2<CarryT>(xy, xz);
                mi.x0 = xa;
                mi.xlast = xz;
                MISC0[par] = mi;
            }
            level_sync(par);
            resolve(par);
            if (g == 0 && lane == 0 && p.input_knots) p.input_knots[sig] = K;
        }

        // =====================================================================================
        // level loop (ITD.py:389-432)
        // =====================================================================================
        int e = 0, stop_kind_v = 0;
        for (;; ++e) {
            const bool last = (e == p.emax);
            const int qn = par ^ 1;                        // buffers of the NEXT level
            OutT *rrow = rot + (long long)e * n;
            OutT *brow = bas ? bas + (long long)e * n : nullptr;
            const bool gen = last || brow != nullptr || !out_vec;
            cnt_lane = 0;
            pend = 0;
            vlast = (CarryT)0;
            CarryT hbl = (CarryT)0, hbr = (CarryT)0;       // B at samples a-1 and b (next level's halo values)
            // knots kb-1 (p2) and kb (p1) relative to the current run start
            int p1t = hk->bt[0], p2t = hk->bt[1];
            CarryT p1x = hk->bx[0], p2x = hk->bx[1];
            int krun = kb;                                 // global index of knot p1
            const unsigned *mk = Mk + par * mwords;
            unsigned *mkn = Mk + qn * mwords;
            const int wend = (u1 - cu0) * SPL;             // end of my mask words (chunk-relative)
            Misc *mo = &MISC0[qn];

            // one unit of samples.  slot: table slot of the segment holding the sample before the lane's first
            auto unit_body = [&](auto edge_c, auto dense_c, auto gen_c, int uu, unsigned fb, int slot) {
                constexpr bool EDGE = decltype(edge_c)::value, DENSE = decltype(dense_c)::value, GEN = decltype(gen_c)::value;
                const int cu = uu - cu0, tl = uu * UNIT + lane * SPL;
                CarryT *xu = Xs + (size_t)cu * UNIT;
                // x of the unit (swizzled chunks) -> registers; save it for the knot-stop trend row
                CarryT xv[SPL];
#pragma unroll
                for (int qc = 0; qc < LCH; ++qc) {
                    const CVec d = *reinterpret_cast<const CVec *>(xu + coff[qc]);
                    *reinterpret_cast<CVec *>(bk + tl + qc * EPC) = d;
                    if constexpr (sizeof(CarryT) == 8) { xv[qc * 2] = d.x; xv[qc * 2 + 1] = d.y; }
                    else { xv[qc * 4] = d.x; xv[qc * 4 + 1] = d.y; xv[qc * 4 + 2] = d.z; xv[qc * 4 + 3] = d.w; }
                }
                CarryT bv[SPL];
                if (DENSE) {
                    const CarryT *xt = XT + slot;
                    const LS *lp = ls + slot;
#pragma unroll
                    for (int j = 0; j < SPL; ++j) {
                        const unsigned bit = (fb >> j) & 1u;
                        xt += bit;
                        lp += bit;
                        const LS q1 = *lp;
                        bv[j] = A::add(q1.L, A::mul(q1.s, A::sub(xv[j], *xt)));      // ITD.py:115-117
                    }
                } else {
                    const CarryT Xk = XT[slot];
                    const LS q1 = ls[slot];
#pragma unroll
                    for (int j = 0; j < SPL; ++j) bv[j] = A::add(q1.L, A::mul(q1.s, A::sub(xv[j], Xk)));
                }
                if (EDGE) {
#pragma unroll
                    for (int j = 0; j < SPL; ++j)
                        if (tl + j >= n - 1) bv[j] = (CarryT)0;                         // ITD.py:112
                }
                // R = x - B (ITD.py:119); the iteration stop's last row is R + B (ITD.py:420)
                if (!EDGE && !GEN) {
#pragma unroll
                    for (int qv = 0; qv < SPL / OPV; ++qv) {
                        OVec o;
                        if constexpr (sizeof(OutT) == 8) {
                            o.x = (OutT)A::sub(xv[qv * 2], bv[qv * 2]); o.y = (OutT)A::sub(xv[qv * 2 + 1], bv[qv * 2 + 1]);
                        } else {
                            o.x = (OutT)A::sub(xv[qv * 4], bv[qv * 4]); o.y = (OutT)A::sub(xv[qv * 4 + 1], bv[qv * 4 + 1]);
                            o.z = (OutT)A::sub(xv[qv * 4 + 2], bv[qv * 4 + 2]); o.w = (OutT)A::sub(xv[qv * 4 + 3], bv[qv * 4 + 3]);
                        }
                        *reinterpret_cast<OVec *>(rrow + tl + qv * OPV) = o;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < SPL; ++j) {
                        if (!EDGE || tl + j < n) {
                            const CarryT r0 = A::sub(xv[j], bv[j]);
                            rrow[tl + j] = (OutT)(last ? A::add(r0, bv[j]) : r0);
                            if (brow) brow[tl + j] = last ? (OutT)0 : (OutT)bv[j];      // ITD.py:424
                        }
                    }
                }
                // carry <- B
#pragma unroll
                for (int qc = 0; qc < LCH; ++qc) {
                    CVec o;
                    if constexpr (sizeof(CarryT) == 8) { o.x = bv[qc * 2]; o.y = bv[qc * 2 + 1]; }
                    else { o.x = bv[qc * 4]; o.y = bv[qc * 4 + 1]; o.z = bv[qc * 4 + 2]; o.w = bv[qc * 4 + 3]; }
                    *reinterpret_cast<CVec *>(xu + coff[qc]) = o;
                }
                // the two end-knot baselines of the next level (ITD.py:101-102 on B)
                if (EDGE) {
                    if (tl == 0) {
                        mo->x0 = bv[0];
                        mo->endl0 = mean2<CarryT>(bv[0], bv[1]);
                    }
#pragma unroll
                    for (int j = 0; j < SPL; ++j) {
                        if (tl + j == n - 2) {
                            mo->endl1 = mean2<CarryT>(bv[j], (CarryT)0);
                            mo->xlast = (CarryT)0;
                        }
                    }
                }
                // extrema of B: the stop test (ITD.py:400-404) and the next level's knots
                emit_flags(edge_c, bv, mkn, uu);
            };

            // per-unit knot counts of my range, one unit per lane (window of 32 units starting at ub)
            int ub = u0;
            auto load_counts = [&](int ubase) -> int {
                const int uu = ubase + lane;
                int c = 0;
                if (uu < u1) {
                    const uint4 *w4 = reinterpret_cast<const uint4 *>(mk + (uu - cu0) * SPL);
#pragma unroll
                    for (int i = 0; i < SPL / 4; ++i) {
                        const uint4 w = w4[i];
                        c += __popc(w.x) + __popc(w.y) + __popc(w.z) + __popc(w.w);
                    }
                }
                return c;
            };
            int cntv = load_counts(ub);

            for (int u = u0; u < u1;) {
                // ---- run = units [u, ue) with at most UNIT knots in total -------------------------
                int ue = u, cntrun = 0;
                while (ue < u1) {
                    if (ue - ub >= 32) {
                        ub = ue;
                        cntv = load_counts(ub);
                    }
                    const int c = __shfl_sync(FULL, cntv, ue - ub);
                    if (ue > u && cntrun + c > UNIT) break;
                    cntrun += c;
                    ++ue;
                }
                // ---- table: slots 0,1 = knots before; 2.. = the run's knots; then three after ------
                if (lane == 0) {
                    TAU[0] = p2t; XT[0] = p2x;
                    TAU[1] = p1t; XT[1] = p1x;
                }
                if (cntrun > 0) {
                    int runpre = 0;
                    for (int uu = u; uu < ue; ++uu) {
                        const int cu = uu - cu0;
                        unsigned fb = (mk[cu * SPL + lw] >> lsh) & FBM;
                        if (!__any_sync(FULL, fb != 0u)) continue;
                        const int c = __popc(fb);
                        int inc = c;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const int t = __shfl_up_sync(FULL, inc, o);
                            if (lane >= o) inc += t;
                        }
                        int idx = 2 + runpre + inc - c;
                        const int trel0 = cu * UNIT + lane * SPL;
                        const CarryT *xu = Xs + (size_t)cu * UNIT;
                        while (fb) {
                            const int j = __ffs(fb) - 1;
                            fb &= fb - 1;
                            TAU[idx] = cu0 * UNIT + trel0 + j;
                            XT[idx] = xu[(lane * LCH + ((j / EPC) ^ xorv)) * EPC + (j % EPC)];
                            ++idx;
                        }
                        runpre += __shfl_sync(FULL, inc, 31);
                    }
                }
                {
                    // look ahead in my own mask for up to three knots after the run
                    int found = 0, ft0 = 0, ft1 = 0, ft2 = 0;
                    for (int wb = (ue - cu0) * SPL; wb < wend && found < 3; wb += 32) {
                        const unsigned wd = (wb + lane < wend) ? mk[wb + lane] : 0u;
                        unsigned nz = __ballot_sync(FULL, wd != 0u);
                        while (nz && found < 3) {
                            const int fl = __ffs(nz) - 1;
                            nz &= nz - 1;
                            unsigned ww = __shfl_sync(FULL, wd, fl);
                            while (ww && found < 3) {
                                const int t = (wb + fl) * 32 + (__ffs(ww) - 1);     // chunk-relative
                                ww &= ww - 1;
                                if (found == 0) ft0 = t; else if (found == 1) ft1 = t; else ft2 = t;
                                ++found;
                            }
                        }
                    }
                    if (lane < 3) {
                        // own-range knots first, then the neighbours' (hk->at), which already end in the virtual end knot
                        int t;
                        CarryT v;
                        if (lane < found) {
                            const int tr = (lane == 0) ? ft0 : (lane == 1) ? ft1 : ft2;
                            t = cu0 * UNIT + tr;
                            v = Xs[swz(tr)];
                        } else {
                            t = hk->at[lane - found];
                            v = hk->ax[lane - found];
                        }
                        TAU[2 + cntrun + lane] = t;
                        XT[2 + cntrun + lane] = v;
                    }
                }
                __syncwarp();
                // ---- knot baseline and slopes, one lane per knot (ITD.py:100-110, :116) ------------
                // slot i <-> global knot index k = krun - 1 + i; L for slots 1..cntrun+3, s for 1..cntrun+2
                for (int i = 1 + lane; i <= cntrun + 3; i += 32) {
                    const int k = krun - 1 + i;
                    CarryT L;
                    if (k <= 0) {
                        L = endl0;
                    } else if (k >= K + 1) {
                        L = endl1;
                    } else {
                        const CarryT w = A::ratio(TAU[i] - TAU[i - 1], TAU[i + 1] - TAU[i - 1]);
                        const CarryT d = A::sub(XT[i + 1], XT[i - 1]);
                        const CarryT qq = A::add(XT[i - 1], A::mul(w, d));
                        L = A::add(A::mul((CarryT)0.5, qq), A::mul((CarryT)0.5, XT[i]));
                    }
                    ls[i].L = L;
                }
                __syncwarp();
                for (int i = 1 + lane; i <= cntrun + 2; i += 32) {
                    const int k = krun - 1 + i;                        // segment [k, k+1)
                    const CarryT den = A::sub(XT[i + 1], XT[i]);
                    ls[i].s = A::div(A::sub(ls[i + 1].L, ls[i].L), den);
                    zero_dx |= (k <= K && den == (CarryT)0);
                }
                __syncwarp();
                // halo B values (the neighbours' samples next to my range), evaluated with my table
                if (u == u0 && a > 0) {
                    const LS q1 = ls[1];
                    hbl = A::add(q1.L, A::mul(q1.s, A::sub(hxl, XT[1])));
                    vlast = hbl;
                }
                if (ue == u1 && b < n) {
                    if (b == n - 1) {
                        hbr = (CarryT)0;                                   // ITD.py:112
                    } else {
                        const int sl = 1 + cntrun + ((TAU[2 + cntrun] == b) ? 1 : 0);
                        const LS q1 = ls[sl];
                        hbr = A::add(q1.L, A::mul(q1.s, A::sub(hxr, XT[sl])));
                    }
                }

                // ---- the samples of the run ----------------------------------------------------------
                int runpre = 0;
                for (int uu = u; uu < ue; ++uu) {
                    const int t0 = uu * UNIT;
                    const bool edge = (uu == 0) || (t0 + UNIT >= n - 1);
                    unsigned fb = 0u;
                    bool dense = false;
                    if (cntrun > 0) {
                        fb = (mk[(uu - cu0) * SPL + lw] >> lsh) & FBM;
                        dense = __any_sync(FULL, fb != 0u);
                    }
                    int slot = 1 + runpre;
                    if (dense) {
                        const int c = __popc(fb);
                        int inc = c;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const int t = __shfl_up_sync(FULL, inc, o);
                            if (lane >= o) inc += t;
                        }
                        slot += inc - c;
                        runpre += __shfl_sync(FULL, inc, 31);
                    }
                    if (edge) unit_body(std::true_type{}, std::true_type{}, std::true_type{}, uu, fb, slot);
                    else if (gen) {
                        if (dense) unit_body(std::false_type{}, std::true_type{}, std::true_type{}, uu, fb, slot);
                        else unit_body(std::false_type{}, std::false_type{}, std::true_type{}, uu, fb, slot);
                    } else {
                        if (dense) unit_body(std::false_type{}, std::true_type{}, std::false_type{}, uu, fb, slot);
                        else unit_body(std::false_type{}, std::false_type{}, std::false_type{}, uu, fb, slot);
                    }
                }
                // ---- the run's last two knots become the next run's "before" knots -------------------
                if (cntrun >= 2) {
                    p2t = TAU[cntrun]; p2x = XT[cntrun];
                    p1t = TAU[cntrun + 1]; p1x = XT[cntrun + 1];
                } else if (cntrun == 1) {
                    p2t = p1t; p2x = p1x;
                    p1t = TAU[2]; p1x = XT[2];
                }
                krun += cntrun;
                u = ue;
                __syncwarp();
            }
            close_range(mkn, hbr);
            hxl = hbl;
            hxr = hbr;
            const int cnt_total = __reduce_add_sync(FULL, cnt_lane);
            publish(qn, cnt_total);
            level_sync(qn);
            par = qn;
            resolve(par);
            if (g == 0 && lane == 0) p.knot_counts[sig * p.rows + e] = K;          // what ITD.py:403 prints
            if (K < p.min_extrema) {                                               // ITD.py:404
                stop_kind_v = kStopKnots;
                break;
            }
            if (last) {                                                            // ITD.py:418
                stop_kind_v = kStopIter;
                break;
            }
        }

        // =====================================================================================
        // stop: trend row, bookkeeping, optional zero tail
        // =====================================================================================
        const int nrows = e + 1;
        if (stop_kind_v == kStopKnots) {
            // the discarded extraction e wrote R_e into row e; the reference returns baselines[e-1] there,
            // i.e. the INPUT of extraction e (zeros when e == 0) (ITD.py:410-411)
            OutT *rrow = rot + (long long)e * n;
            OutT *brow = (bas && (p.opts & kOptZeroTail)) ? bas + (long long)e * n : nullptr;
            for (int t = a + lane; t < b; t += 32) {
                rrow[t] = (e == 0) ? (OutT)0 : (OutT)bk[t];
                if (brow) brow[t] = (OutT)0;
            }
        }
        if (p.opts & kOptZeroTail) {
            for (int r = nrows; r < p.rows; ++r) {
                OutT *rrow = rot + (long long)r * n;
                OutT *brow = bas ? bas + (long long)r * n : nullptr;
                for (int t = a + lane; t < b; t += 32) {
                    rrow[t] = (OutT)0;
                    if (brow) brow[t] = (OutT)0;
                }
            }
        }
        if (g == 0 && lane == 0) {
            p.n_rows[sig] = nrows;
            p.stop_kind[sig] = stop_kind_v;
        }
        const unsigned stbits = (__any_sync(FULL, bad) ? kStNonFinite : 0) | (__any_sync(FULL, zero_dx) ? kStZeroDx : 0);
        if (stbits && lane == 0) atomicOr(p.status + sig, (int)stbits);
    }
    cluster.sync();      // no CTA may exit while a neighbour can still read its shared memory
}

}  // namespace pyitd
