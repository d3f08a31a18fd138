// itd_regres.cuh -- the whole decomposition of a signal in ONE kernel with the signal resident in REGISTERS.
//
// One thread-block cluster (1..8 CTAs of 16 warps) owns one signal at a time.  Every lane keeps SPL
// consecutive samples of the carry X_e in registers for all levels (a warp = one UNIT of 32*SPL samples,
// a cluster of 8 CTAs = 65 536 samples), so per level the only HBM traffic is the rotation row going out:
// x is read once, every output row is written once.  The level loop, the stop test (ITD.py:400-404, :418)
// and the trend-row fix-up (ITD.py:410-411) run inside the kernel: one launch per batch, no host sync.
//
// Per level and warp
//   1. the knots of the warp's samples (flag bits held in a register since the previous level) are already
//      enumerated in the warp's shared-memory table; the two knots before and the three after the warp's
//      span come from the neighbours' published summaries (own CTA: shared memory, other CTAs: DSMEM);
//   2. knot baseline L_k (ITD.py:100-110) and segment slopes (ITD.py:116), one lane per knot;
//   3. per sample B = L_k + s_k (x - X_k), R = x - B (ITD.py:115-119); R -> HBM (128-bit stores), B replaces
//      x in the registers;
//   4. extrema of B with the neighbours in registers (= the stop test = the next level's knots), their
//      enumeration into the table, the summary {count, first three, last two knots}, one cluster barrier.
//
// X_e must survive until the stop test on B_e is known (the knot stop returns X_e as the trend row,
// ITD.py:410-411): each level first saves its input to a per-cluster backup area (L2-resident scratch).
//
// fp64 arithmetic uses the unfused intrinsics of itd_kernels.cuh in the reference's operation order.
#pragma once

#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "itd_kernels.cuh"
#include "itd_resident.cuh"

namespace pyitd {

struct RegResParams {
    const void *x;              // [S, n] input type
    void *rot;                  // [S, rows, n] output type
    void *bas;                  // [S, rows, n] output type or null
    void *backup;               // [clusters, backup_stride] carry type
    long long out_sig_stride;   // rows * n
    long long backup_stride;    // CL * WARPS * UNIT
    int *n_rows, *knot_counts, *input_knots, *stop_kind, *status;
    long long S;
    int n;
    int emax, rows, min_extrema;
    unsigned opts;
};

template <typename CarryT, int WARPS, int SPL>
struct RegResGeom {
    static constexpr int UNIT = 32 * SPL;
    static constexpr int CAP = UNIT + 8;                       // knot-table capacity per warp
    // per warp: {L, s|tau}[CAP], X[CAP], halo knots
    static constexpr size_t tab_core_bytes = ((size_t)CAP * (sizeof(KnotLS<CarryT>) + sizeof(CarryT)) + 15) & ~(size_t)15;
    static constexpr size_t tab_bytes_per_warp = tab_core_bytes + sizeof(HaloKnots<CarryT>);
    static constexpr size_t off_ws = tab_bytes_per_warp * WARPS;
    static constexpr size_t off_cs = off_ws + 2 * WARPS * sizeof(KnotSummary<CarryT>);
    static constexpr size_t off_misc = off_cs + 2 * sizeof(KnotSummary<CarryT>);
    static constexpr size_t smem_bytes = off_misc + 2 * sizeof(LevelMisc<CarryT>);
};

// knot flags of the lane's SPL values; vl = left neighbour of v[0]; lane 31's right neighbour is vr
template <int SPL, typename CarryT>
__device__ __forceinline__ unsigned lane_flags_lr(const CarryT (&v)[SPL], CarryT vl, CarryT vr, int lane) {
    unsigned lt = 0u, gt = 0u;
    cmp_bits(vl, v[0], 1u, lt, gt);
#pragma unroll
    for (int j = 1; j < SPL; ++j) cmp_bits(v[j - 1], v[j], 1u << j, lt, gt);
    unsigned nx = __shfl_down_sync(0xffffffffu, (lt & 1u) | ((gt & 1u) << 1), 1);
    if (lane == 31) nx = (v[SPL - 1] < vr ? 1u : 0u) | (v[SPL - 1] > vr ? 2u : 0u);
    const unsigned lte = lt | ((nx & 1u) << SPL), gte = gt | ((nx >> 1) << SPL);
    // valley: !(x[t-1] < x[t]) && x[t] < x[t+1];  peak: !(x[t-1] > x[t]) && x[t] > x[t+1]   (ITD.py:59 on x and -x)
    return (((~lt) & (lte >> 1)) | ((~gt) & (gte >> 1))) & ((1u << SPL) - 1u);
}

template <typename InT, typename CarryT, typename OutT, int WARPS, int SPL>
__global__ void __launch_bounds__(WARPS * 32, 1) regres_kernel(const RegResParams p) {
    using A = Arith<CarryT>;
    using G = RegResGeom<CarryT, WARPS, SPL>;
    using Summary = KnotSummary<CarryT>;
    using Misc = LevelMisc<CarryT>;
    using LS = KnotLS<CarryT>;
    using Halo = HaloKnots<CarryT>;
    using CVec = typename std::conditional<sizeof(CarryT) == 8, double2, float4>::type;
    using OVec = typename std::conditional<sizeof(OutT) == 8, double2, float4>::type;
    constexpr int UNIT = G::UNIT, CAP = G::CAP;
    constexpr int EPC = 16 / (int)sizeof(CarryT);          // carry elements per 16 bytes
    constexpr int OPV = 16 / (int)sizeof(OutT);            // output elements per 16-byte store
    constexpr unsigned FBM = (SPL == 32) ? 0xffffffffu : ((1u << SPL) - 1u);
    constexpr unsigned FULL = 0xffffffffu;
    static_assert(SPL == 8 || SPL == 16, "SPL must be 8 or 16");

    extern __shared__ __align__(128) unsigned char smem_rr[];
    cg::cluster_group cluster = cg::this_cluster();
    const int CL = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int cid = (int)(blockIdx.x / CL), ncl = (int)(gridDim.x / CL);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = rank * WARPS + warp;
    const int n = p.n;

    // ---- this warp's unit and this lane's samples ------------------------------------------------
    const int a = g * UNIT;
    const bool have = a < n;
    const int b = have ? min(a + UNIT, n) : a;                // samples [a, b)
    const int tl = a + lane * SPL;                            // the lane's first sample
    const bool edge = (a == 0) || (a + UNIT >= n - 1);        // holds sample 0, n-2, n-1 or padding
    unsigned vmask = 0u, zmask = 0u, smask = 0u;              // flaggable / forced-zero / stored samples of the lane
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        const int t = tl + j;
        if (t >= 1 && t <= n - 2) vmask |= 1u << j;
        if (t >= n - 1) zmask |= 1u << j;
        if (t < n) smask |= 1u << j;
    }

    unsigned char *tab = smem_rr + (size_t)warp * G::tab_bytes_per_warp;
    LS *ls = reinterpret_cast<LS *>(tab);
    CarryT *XT = reinterpret_cast<CarryT *>(tab + (size_t)CAP * sizeof(LS));
    Halo *hk = reinterpret_cast<Halo *>(tab + G::tab_core_bytes);
    Summary *WS = reinterpret_cast<Summary *>(smem_rr + G::off_ws);       // [2][WARPS]
    Summary *CS = reinterpret_cast<Summary *>(smem_rr + G::off_cs);       // [2]
    Misc *MISC = reinterpret_cast<Misc *>(smem_rr + G::off_misc);         // [2], the copy in CTA 0 is the live one
    Misc *MISC0 = cluster.map_shared_rank(MISC, 0);
    // knot positions live in the slope slot until the slopes are computed
    auto tau = [&](int i) -> int & { return *reinterpret_cast<int *>(&ls[i].s); };

    int par = 0;                       // parity of the summary buffers holding the CURRENT level's knots
    CarryT *bk = reinterpret_cast<CarryT *>(p.backup) + (long long)cid * p.backup_stride;

    for (long long sig = cid; sig < p.S; sig += ncl) {
        const InT *x = reinterpret_cast<const InT *>(p.x) + sig * n;
        OutT *rot = reinterpret_cast<OutT *>(p.rot) + sig * p.out_sig_stride;
        OutT *bas = p.bas ? reinterpret_cast<OutT *>(p.bas) + sig * p.out_sig_stride : nullptr;
        const bool in_vec = ((reinterpret_cast<uintptr_t>(x) | ((size_t)n * sizeof(InT))) & 15) == 0;
        const bool out_vec = ((reinterpret_cast<uintptr_t>(rot) | ((size_t)n * sizeof(OutT))) & 15) == 0 &&
                             (!bas || (reinterpret_cast<uintptr_t>(bas) & 15) == 0);
        bool bad = false, zero_dx = false;
        par ^= 1;        // slow warps may still be reading the previous signal's last summaries (buffer par)

        // =====================================================================================
        // helpers
        // =====================================================================================
        // the lane's flagged samples -> table slots 2.. (position, value); returns the warp's knot count
        // and the lane's exclusive prefix.  v are the lane's values (registers).
        auto enumerate = [&](unsigned f, const CarryT (&v)[SPL], int &excl) -> int {
            const int c = __popc(f);
            int inc = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(FULL, inc, o);
                if (lane >= o) inc += t;
            }
            excl = inc - c;
            const int total = __shfl_sync(FULL, inc, 31);
            if (total > 0) {
                int idx = 2 + excl;
#pragma unroll
                for (int j = 0; j < SPL; ++j) {
                    if ((f >> j) & 1u) {
                        tau(idx) = tl + j;
                        XT[idx] = v[j];
                        ++idx;
                    }
                }
            }
            __syncwarp();
            return total;
        };
        // summary of the warp's knots (already in table slots 2..cnt+1) into buffer q
        auto publish = [&](int q, int cnt) {
            if (lane == 0) {
                Summary *me = &WS[q * WARPS + warp];
                me->cnt = cnt;
                if (cnt > 0) {
                    me->tF[0] = tau(2); me->xF[0] = XT[2];
                    me->tL[0] = tau(1 + cnt); me->xL[0] = XT[1 + cnt];
                    if (cnt > 1) {
                        me->tF[1] = tau(3); me->xF[1] = XT[3];
                        me->tL[1] = tau(cnt); me->xL[1] = XT[cnt];
                    }
                    if (cnt > 2) { me->tF[2] = tau(4); me->xF[2] = XT[4]; }
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
                    Summary *d = &CS[q];
                    const unsigned m = __ballot_sync(FULL, c > 0);
                    int found = 0;
                    unsigned mm = m;
                    while (mm && found < 3) {                     // first three knots of the CTA
                        const int j = __ffs(mm) - 1;
                        mm &= mm - 1;
                        const int take = min(min(__shfl_sync(FULL, c, j), 3), 3 - found);
                        if (lane == j)
                            for (int i = 0; i < take; ++i) { d->tF[found + i] = src[j].tF[i]; d->xF[found + i] = src[j].xF[i]; }
                        found += take;
                    }
                    found = 0;
                    mm = m;
                    while (mm && found < 2) {                     // last two knots of the CTA
                        const int j = 31 - __clz(mm);
                        mm &= ~(1u << j);
                        const int take = min(min(__shfl_sync(FULL, c, j), 2), 2 - found);
                        if (lane == j)
                            for (int i = 0; i < take; ++i) { d->tL[found + i] = src[j].tL[i]; d->xL[found + i] = src[j].xL[i]; }
                        found += take;
                    }
                    if (lane == 0) d->cnt = tot;
                }
            }
            cluster.sync();
        };
        // neighbourhood of this warp's span at the current level: scalars in registers, knots in hk
        int K = 0, kb = 0;
        CarryT endl0 = 0, endl1 = 0;
        auto resolve = [&](int q, int mycnt) {
            // entry list in sample order: CTAs before mine (aggregates), my CTA's warps, CTAs after mine
            const int ne = CL - 1 + WARPS, me = rank + warp;
            const Summary *e = nullptr;
            if (lane < rank) e = cluster.map_shared_rank(&CS[q], lane);
            else if (lane < rank + WARPS) e = &WS[q * WARPS + (lane - rank)];
            else if (lane < ne) e = cluster.map_shared_rank(&CS[q], lane - WARPS + 1);
            const int c = e ? e->cnt : 0;
            const Misc mi = MISC0[q];
            endl0 = mi.endl0; endl1 = mi.endl1;
            K = __reduce_add_sync(FULL, c);
            kb = __reduce_add_sync(FULL, (lane < me) ? c : 0);
            const unsigned nzm = __ballot_sync(FULL, c > 0);
            // two nearest real knots before the span: the owning lanes write them
            unsigned m = nzm & ((1u << me) - 1u);
            if (m) {
                const int j = 31 - __clz(m);
                m &= ~(1u << j);
                const int cj = __shfl_sync(FULL, c, j);
                if (lane == j) {
                    hk->bt[0] = e->tL[0]; hk->bx[0] = e->xL[0];
                    if (c >= 2) { hk->bt[1] = e->tL[1]; hk->bx[1] = e->xL[1]; }
                }
                if (cj < 2 && m) {
                    const int j2 = 31 - __clz(m);
                    if (lane == j2) { hk->bt[1] = e->tL[0]; hk->bx[1] = e->xL[0]; }
                }
            }
            // three nearest real knots after the span
            m = (me >= 31) ? 0u : (nzm & ~((2u << me) - 1u));
            int na = 0;
#pragma unroll
            for (int it = 0; it < 3; ++it) {
                if (m && na < 3) {
                    const int j = __ffs(m) - 1;
                    m &= m - 1;
                    const int take = min(min(__shfl_sync(FULL, c, j), 3), 3 - na);
                    if (lane == j)
                        for (int i = 0; i < take; ++i) { hk->at[na + i] = e->tF[i]; hk->ax[na + i] = e->xF[i]; }
                    na += take;
                }
            }
            __syncwarp();
            if (lane == 0) {
                // knot kb (nearest before) and kb-1: real, or the virtual start knot (tau 0, x[0]), or unused
                if (kb < 1) { hk->bt[0] = 0; hk->bx[0] = mi.x0; }
                if (kb < 2) { hk->bt[1] = 0; hk->bx[1] = mi.x0; }
                // knots ka+1.. (ka = kb + mycnt): real while <= K, then the virtual end knot (tau n-1, x[n-1])
                const int ka = kb + mycnt;
#pragma unroll
                for (int i = 0; i < 3; ++i)
                    if (ka + 1 + i > K) { hk->at[i] = n - 1; hk->ax[i] = mi.xlast; }
            }
            __syncwarp();
        };

        // =====================================================================================
        // 0. load the unit into registers, detect the extrema of the input (ITD.py:87-98)
        // =====================================================================================
        CarryT xr[SPL];                   // the lane's samples of X_e, for the whole decomposition
        CarryT hxl = (CarryT)0, hxr = (CarryT)0;      // values of samples a-1 and b at the current level
        unsigned fb = 0u;                 // knot flags of the lane's samples at the current level
        int excl = 0, mycnt = 0;          // knots of the warp before the lane's first sample / in the warp
        {
            if (have && !edge && in_vec) {
                constexpr int IPV = 16 / (int)sizeof(InT);
#pragma unroll
                for (int qv = 0; qv < SPL / IPV; ++qv) {
                    if constexpr (sizeof(InT) == 8) {
                        const double2 d = __ldg(reinterpret_cast<const double2 *>(x + tl) + qv);
                        xr[qv * 2] = (CarryT)d.x; xr[qv * 2 + 1] = (CarryT)d.y;
                    } else {
                        const float4 d = __ldg(reinterpret_cast<const float4 *>(x + tl) + qv);
                        xr[qv * 4] = (CarryT)d.x; xr[qv * 4 + 1] = (CarryT)d.y; xr[qv * 4 + 2] = (CarryT)d.z; xr[qv * 4 + 3] = (CarryT)d.w;
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < SPL; ++j) xr[j] = (tl + j < n) ? (CarryT)__ldg(x + tl + j) : (CarryT)0;
            }
#pragma unroll
            for (int j = 0; j < SPL; ++j) bad |= !isfinite(xr[j]);
            if (have) {
                if (a > 0) hxl = (CarryT)__ldg(x + a - 1);
                if (b < n) hxr = (CarryT)__ldg(x + b);
            }
            CarryT vl = __shfl_up_sync(FULL, xr[SPL - 1], 1);
            if (lane == 0) vl = hxl;
            fb = lane_flags_lr<SPL, CarryT>(xr, vl, hxr, lane) & vmask;
            mycnt = enumerate(fb, xr, excl);
            publish(par, mycnt);
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
            resolve(par, mycnt);
            if (g == 0 && lane == 0 && p.input_knots) p.input_knots[sig] = K;
        }

        // =====================================================================================
        // level loop (ITD.py:389-432)
        // =====================================================================================
        int e = 0, stop_kind_v = 0;
        for (;; ++e) {
            const bool last = (e == p.emax);
            const int qn = par ^ 1;                        // summary buffers of the NEXT level
            OutT *rrow = rot + (long long)e * n;
            OutT *brow = bas ? bas + (long long)e * n : nullptr;
            const bool gen = last || brow != nullptr || !out_vec;
            Misc *mo = &MISC0[qn];
            CarryT hbl = (CarryT)0, hbr = (CarryT)0;       // B at samples a-1 and b (next level's halo values)

            // ---- table: slots 0,1 = knots before; 2..mycnt+1 = the warp's knots; then three after ----
            if (lane == 0) {
                tau(0) = hk->bt[1]; XT[0] = hk->bx[1];
                tau(1) = hk->bt[0]; XT[1] = hk->bx[0];
            }
            if (lane < 3) {
                tau(2 + mycnt + lane) = hk->at[lane];
                XT[2 + mycnt + lane] = hk->ax[lane];
            }
            const bool right_is_knot = (hk->at[0] == b);
            __syncwarp();
            // ---- knot baseline and slopes, one lane per knot (ITD.py:100-110, :116) -----------------
            // slot i <-> global knot index k = kb - 1 + i; L for slots 1..mycnt+3, s for 1..mycnt+2
            for (int i = 1 + lane; i <= mycnt + 3; i += 32) {
                const int k = kb - 1 + i;
                CarryT L;
                if (k <= 0) {
                    L = endl0;
                } else if (k >= K + 1) {
                    L = endl1;
                } else {
                    const CarryT w = A::ratio(tau(i) - tau(i - 1), tau(i + 1) - tau(i - 1));
                    const CarryT d = A::sub(XT[i + 1], XT[i - 1]);
                    const CarryT qq = A::add(XT[i - 1], A::mul(w, d));
                    L = A::add(A::mul((CarryT)0.5, qq), A::mul((CarryT)0.5, XT[i]));
                }
                ls[i].L = L;
            }
            __syncwarp();
            for (int i = 1 + lane; i <= mycnt + 2; i += 32) {
                const int k = kb - 1 + i;                          // segment [k, k+1)
                const CarryT den = A::sub(XT[i + 1], XT[i]);
                ls[i].s = A::div(A::sub(ls[i + 1].L, ls[i].L), den);
                zero_dx |= (k <= K && den == (CarryT)0);
            }
            __syncwarp();
            // halo B values (the neighbours' samples next to my span), evaluated with my table
            if (have) {
                if (a > 0) {
                    const LS q1 = ls[1];
                    hbl = A::add(q1.L, A::mul(q1.s, A::sub(hxl, XT[1])));
                }
                if (b < n && b != n - 1) {                         // B[n-1] = 0 (ITD.py:112)
                    const int sl = 1 + mycnt + (right_is_knot ? 1 : 0);
                    const LS q1 = ls[sl];
                    hbr = A::add(q1.L, A::mul(q1.s, A::sub(hxr, XT[sl])));
                }
            }

            // ---- the samples: B replaces x in the registers, R goes out ------------------------------
            auto sample_pass = [&](auto edge_c, auto dense_c, auto gen_c) {
                constexpr bool EDGE = decltype(edge_c)::value, DENSE = decltype(dense_c)::value, GEN = decltype(gen_c)::value;
                // save x for the knot-stop trend row
#pragma unroll
                for (int qc = 0; qc < SPL / EPC; ++qc) {
                    CVec o;
                    if constexpr (sizeof(CarryT) == 8) { o.x = xr[qc * 2]; o.y = xr[qc * 2 + 1]; }
                    else { o.x = xr[qc * 4]; o.y = xr[qc * 4 + 1]; o.z = xr[qc * 4 + 2]; o.w = xr[qc * 4 + 3]; }
                    *reinterpret_cast<CVec *>(bk + tl + qc * EPC) = o;
                }
                const CarryT *xt = XT + 1 + (DENSE ? excl : 0);
                const LS *lp = ls + 1 + (DENSE ? excl : 0);
                CarryT Xk = *xt;
                LS q1 = *lp;
                CarryT rv[OPV];
#pragma unroll
                for (int j = 0; j < SPL; ++j) {
                    if (DENSE) {
                        const unsigned bit = (fb >> j) & 1u;
                        xt += bit;
                        lp += bit;
                        Xk = *xt;
                        q1 = *lp;
                    }
                    CarryT bb = A::add(q1.L, A::mul(q1.s, A::sub(xr[j], Xk)));      // ITD.py:115-117
                    if (EDGE && ((zmask >> j) & 1u)) bb = (CarryT)0;               // ITD.py:112
                    const CarryT r0 = A::sub(xr[j], bb);                           // ITD.py:119
                    xr[j] = bb;
                    if (!EDGE && !GEN) {
                        rv[j % OPV] = r0;
                        if (j % OPV == OPV - 1) {
                            OVec o;
                            if constexpr (sizeof(OutT) == 8) { o.x = (OutT)rv[0]; o.y = (OutT)rv[1]; }
                            else { o.x = (OutT)rv[0]; o.y = (OutT)rv[1]; o.z = (OutT)rv[2]; o.w = (OutT)rv[3]; }
                            *reinterpret_cast<OVec *>(rrow + tl + j - (OPV - 1)) = o;
                        }
                    } else if (!EDGE || ((smask >> j) & 1u)) {
                        rrow[tl + j] = (OutT)(last ? A::add(r0, bb) : r0);          // ITD.py:420
                        if (brow) brow[tl + j] = last ? (OutT)0 : (OutT)bb;        // ITD.py:424
                    }
                }
                // the two end-knot baselines of the next level (ITD.py:101-102 on B)
                if (EDGE) {
                    if (tl == 0) {
                        mo->x0 = xr[0];
                        mo->endl0 = mean2<CarryT>(xr[0], xr[1]);
                    }
#pragma unroll
                    for (int j = 0; j < SPL; ++j) {
                        if (tl + j == n - 2) {
                            mo->endl1 = mean2<CarryT>(xr[j], (CarryT)0);
                            mo->xlast = (CarryT)0;
                        }
                    }
                }
            };
            if (have) {
                const bool dense = mycnt > 0;
                if (edge) sample_pass(std::true_type{}, std::true_type{}, std::true_type{});
                else if (gen) {
                    if (dense) sample_pass(std::false_type{}, std::true_type{}, std::true_type{});
                    else sample_pass(std::false_type{}, std::false_type{}, std::true_type{});
                } else {
                    if (dense) sample_pass(std::false_type{}, std::true_type{}, std::false_type{});
                    else sample_pass(std::false_type{}, std::false_type{}, std::false_type{});
                }
            }
            __syncwarp();       // every lane is done with the table

            // ---- extrema of B: the stop test (ITD.py:400-404) and the next level's knots -------------
            {
                CarryT vl = __shfl_up_sync(FULL, xr[SPL - 1], 1);
                if (lane == 0) vl = hbl;
                fb = lane_flags_lr<SPL, CarryT>(xr, vl, hbr, lane) & vmask;
            }
            hxl = hbl;
            hxr = hbr;
            mycnt = enumerate(fb, xr, excl);
            publish(qn, mycnt);
            level_sync(qn);
            par = qn;
            resolve(par, mycnt);
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
