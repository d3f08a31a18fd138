// itd_stream.cuh -- the batched-channel kernels: ONE CTA PER SIGNAL, tiles walked in order.
//
// Same arithmetic and the same HBM data structures as the look-back kernels (itd_kernels.cuh); the
// difference is how a signal is moved through the SM:
//
//   * each tile's samples, knot-flag words and knot-table slice arrive in a shared-memory ring by
//     TMA bulk copies (cp.async.bulk + mbarrier complete_tx) issued STAGES tiles ahead of the math
//     by one elected thread right after the per-tile barrier (which also proves the stage free);
//   * eight consumer warps each own a contiguous 32*ITEMS-sample span of the tile.  A warp derives
//     its segment ids from the stored flag words (no stencil re-run, no block scan), evaluates the
//     knot baseline / slopes for exactly the knots its span touches (ITD.py:106-110,116) in warp-
//     private scratch, evaluates B and R (ITD.py:115-119), streams them out, and finds the next
//     level's knots from B with shuffles + ballots;
//   * because the tiles of a signal are visited in order by one CTA, the running knot count is a
//     register: no look-back chain, and exactly ONE block barrier per tile (the exchange of the
//     per-warp new-knot counts).
//
// scan_stream_kernel is the same pipeline for the very first pass (extrema of the raw input).
//
// Used when the batch has enough signals to fill the GPU with one CTA each; few long signals use
// the multi-CTA look-back kernels instead.
#pragma once

#include <type_traits>

#include "itd_kernels.cuh"

namespace pyitd {

// ---------------------------------------------------------------------------------------------
// mbarrier / TMA-bulk PTX wrappers (all take 32-bit shared-window addresses)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) {
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// try_wait itself suspends the thread for a bounded time; after a few failed rounds back off with nanosleep so that a
// long wait (a TMA slice behind a busy memory system) does not burn issue slots.  -DPYITD_DEBUG_TRAP turns a wait that
// never ends (a descriptor or byte-count bug) into a trap instead of a hang.
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    if (mbar_try_wait(bar, parity)) return;
    unsigned spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > 16u) __nanosleep(64);
#ifdef PYITD_DEBUG_TRAP
        if (spins > (1u << 24)) __trap();
#endif
    }
}
// global -> shared bulk copy; dst, src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void tma_load_1d(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void named_barrier_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

template <typename T>
__device__ __forceinline__ T shfl_idx(T v, int src) {
    return __shfl_sync(0xffffffffu, v, src);
}

// ---------------------------------------------------------------------------------------------
// shared-memory layout
// ---------------------------------------------------------------------------------------------
template <typename InT, typename CarryT, int WARPS, int ITEMS, int STAGES, bool WITH_KNOTS>
struct StreamSmem {
    static constexpr int T = WARPS * 32 * ITEMS;
    static constexpr int KC = WITH_KNOTS ? T + 16 : 4;   // knot slice capacity (T + 5, start aligned down to 4)
    static constexpr int SPAN = 32 * ITEMS;
    static constexpr int SC = WITH_KNOTS ? SPAN + 8 : 1; // per-warp knot scratch
    static constexpr int MAX_TILES = 1024;
    struct alignas(2 * sizeof(CarryT)) LS {
        CarryT L, s;                             // knot baseline L_k and slope of segment [k, k+1)
    };
    struct Stage {
        alignas(16) InT x[T];
        alignas(16) unsigned mask[(T / 32 + 3) & ~3];
        alignas(16) int tau[KC];
        alignas(16) CarryT xk[KC];
        alignas(16) LS lsg[WITH_KNOTS ? kLsStage : 1];   // slice of the knot_ls_kernel table (tiles with few knots)
        int ls_mode;                                      // 1: lsg holds {L, slope} of knots kb .. kb + cnt + 1
    };
    Stage stage[STAGES];
    alignas(8) unsigned long long full[STAGES];
    LS ls[WARPS][SC];
    int tbase[WITH_KNOTS ? MAX_TILES + 1 : 1];
    int cnt[2][WARPS];
    CarryT carry_b[2];                           // value of the previous tile's last sample (by tile parity)
    CarryT endl[2];
};
constexpr int kStreamMaxTiles = 1024;

// ---------------------------------------------------------------------------------------------
// extrema of a span held in registers (striped: v[r] is sample span0 + r*32 + lane).
// fw[r] = flag word of samples [span0 + 32 r, +32); returns their total.  vleft / vright are the
// samples just outside the span.  EDGE applies the 1 <= t <= n-2 rule (ITD.py:70-73).
// ---------------------------------------------------------------------------------------------
template <bool EDGE, int ITEMS, typename CarryT>
__device__ __forceinline__ int span_extrema(const CarryT (&v)[ITEMS], CarryT vleft, CarryT vright, int lane,
                                            int tspan, int n, unsigned (&fw)[ITEMS]) {
    const CarryT v0 = shfl_idx(v[0], 0);
    unsigned lt_in = (vleft < v0) ? 1u : 0u;
    unsigned gt_in = (vleft > v0) ? 1u : 0u;
    int total = 0;
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
        // right neighbour: lane+1 of the same round; lane 31 takes lane 0 of the next round
        const CarryT give = (lane == 0) ? ((r + 1 < ITEMS) ? v[(r + 1 < ITEMS) ? r + 1 : r] : vright) : v[r];
        const CarryT nx = shfl_idx(give, (lane + 1) & 31);
        const unsigned LT = __ballot_sync(0xffffffffu, v[r] < nx);
        const unsigned GT = __ballot_sync(0xffffffffu, v[r] > nx);
        // valley: !(v[i-1] < v[i]) && v[i] < v[i+1];  peak: !(v[i-1] > v[i]) && v[i] > v[i+1]
        unsigned f = (~((LT << 1) | lt_in) & LT) | (~((GT << 1) | gt_in) & GT);
        lt_in = LT >> 31;
        gt_in = GT >> 31;
        if (EDGE) {
            const int tw = tspan + r * 32;                            // global index of bit 0
            if (tw == 0) f &= ~1u;
            const int lastbit = n - 2 - tw;                           // highest valid bit
            f = (lastbit < 0) ? 0u : ((lastbit >= 31) ? f : (f & (0xffffffffu >> (31 - lastbit))));
        }
        fw[r] = f;
        total += __popc(f);
    }
    return total;
}

// the one block barrier per tile: exchange per-warp counts, then write the compacted knots
template <int WARPS, int ITEMS, typename CarryT>
__device__ __forceinline__ void compact_knots(int (&cnt)[2][WARPS], int i, int warp, int lane, int newc,
                                              const unsigned (&fw)[ITEMS], const CarryT (&v)[ITEMS],
                                              int tspan, int &run_total, int *ntau, CarryT *nxk, int *ntbase) {
    if (lane == 0) cnt[i & 1][warp] = newc;
    named_barrier_sync(1, WARPS * 32);
    const int c = (lane < WARPS) ? cnt[i & 1][lane] : 0;
    const int tot = __reduce_add_sync(0xffffffffu, c);
    int pre = run_total + __reduce_add_sync(0xffffffffu, (lane < warp) ? c : 0);
    if (warp == 0 && lane == 0) ntbase[i] = run_total;
    run_total += tot;
    const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
        if ((fw[r] >> lane) & 1u) {
            const int rank = pre + __popc(fw[r] & lt_mask);
            ntau[1 + rank] = tspan + r * 32 + lane;
            nxk[1 + rank] = v[r];
        }
        pre += __popc(fw[r]);
    }
}

// vectorised row copy / fill used by the trend-row fix-up
template <typename OutT, typename CarryT>
__device__ __forceinline__ void copy_row(OutT *dst, const CarryT *src, int n, bool zero) {
    constexpr int U = 8;                                  // loads in flight per thread (a fix-up block copies 512 KB)
    for (int t = threadIdx.x; t < n; t += blockDim.x * U) {
        CarryT a[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int tt = t + u * blockDim.x;
            a[u] = (!zero && tt < n) ? src[tt] : (CarryT)0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int tt = t + u * blockDim.x;
            if (tt < n) dst[tt] = (OutT)a[u];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// level_stream_kernel
// ---------------------------------------------------------------------------------------------
template <typename InT, typename CarryT, typename OutT, int WARPS, int ITEMS, int STAGES, bool LAST, bool BAS>
__global__ void __launch_bounds__(WARPS * 32, 3) level_stream_kernel(const LevelParams p) {
    using A = Arith<CarryT>;
    using Smem = StreamSmem<InT, CarryT, WARPS, ITEMS, STAGES, true>;
    using LS = typename Smem::LS;
    constexpr int T = Smem::T;
    constexpr int SPAN = Smem::SPAN;
    static_assert(T / 32 <= 32 && ITEMS == 4, "one flag word per lane, four words per warp (LDS.128)");
    extern __shared__ __align__(128) unsigned char smem_stream_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_stream_raw);
    unsigned sbase = smem_u32(smem_stream_raw);
    asm volatile("" : "+r"(sbase));                       // keep it in a register (no S2R re-derivation)
    const unsigned full0 = sbase + (unsigned)offsetof(Smem, full);

    const int sig = blockIdx.x + p.sig0;
    const int n = p.n, e = p.e, tiles = p.tiles;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long row_off = (long long)sig * p.out_sig_stride;

    // ---- signals that already stopped (same rules as level_kernel) ---------------------------
    const int se = p.stop_e[sig];
    if (e > se) {
        OutT *rot = reinterpret_cast<OutT *>(p.rot) + row_off;
        OutT *bas = BAS ? reinterpret_cast<OutT *>(p.bas) + row_off : nullptr;
        const CarryT *src = reinterpret_cast<const CarryT *>(p.fix_src) + (long long)sig * n;
        if (e == se + 1 && p.stop_kind[sig] == kStopKnots) {
            // the discarded extraction `se` wrote R_se into row se; the reference returns
            // baselines[se-1] there, i.e. the INPUT of that extraction (zeros when se == 0)
            copy_row(rot + (long long)se * n, src, n, se == 0);
            if (BAS && (p.opts & kOptZeroTail)) copy_row(bas + (long long)se * n, src, n, true);
        }
        if ((p.opts & kOptZeroTail) && e < p.rows) {
            copy_row(rot + (long long)e * n, src, n, true);
            if (BAS) copy_row(bas + (long long)e * n, src, n, true);
        }
        return;
    }
    if (e > p.emax) return;

    // ---- prologue ---------------------------------------------------------------------------
    const int K = p.cur.kcount[sig];
    {
        const int *gtb = p.cur.tbase + (long long)sig * (tiles + 1);
        for (int i = tid; i <= tiles; i += blockDim.x) sm.tbase[i] = gtb[i];
        if (tid == 0) {
            const CarryT *gendl = reinterpret_cast<const CarryT *>(p.cur.endl) + 2ll * sig;
            sm.endl[0] = gendl[0];
            sm.endl[1] = gendl[1];
            sm.carry_b[0] = sm.carry_b[1] = (CarryT)0;
            for (int s = 0; s < STAGES; ++s) mbar_init(full0 + 8 * s, 1);
            mbar_fence_init();
        }
    }
    __syncthreads();

    // TMA bulk loads of tile i into stage i % STAGES (one thread)
    const InT *x = reinterpret_cast<const InT *>(p.in) + (long long)sig * n;
    const int *gtau = p.cur.tau + (long long)sig * p.cur.kstride;
    const CarryT *gxk = reinterpret_cast<const CarryT *>(p.cur.xk) + (long long)sig * p.cur.kstride;
    const unsigned *gmask_in = p.cur.mask + (long long)sig * p.cur.mstride;
    const LS *gls = reinterpret_cast<const LS *>(p.ls) + (long long)sig * (p.lscap + 4);
    const bool ls_ok = (p.ls != nullptr) && K <= p.lscap;
    auto issue_tile = [&](const int i) {
        const int s = i % STAGES;
        const unsigned st = sbase + (unsigned)(offsetof(Smem, stage) + (size_t)s * sizeof(typename Smem::Stage));
        const int t0 = i * T;
        const int len = min(T, n - t0);
        const int kb = sm.tbase[i], cnt = sm.tbase[i + 1] - kb;
        const int lo = max(kb - 1, 0) & ~3;
        const int hi = min(kb + cnt + 3, K + 1);
        const int nk = (hi - lo + 1 + 3) & ~3;
        const unsigned bx = (unsigned)(len * sizeof(InT));
        const unsigned bm = (unsigned)((((len + 31) / 32 + 3) & ~3) * sizeof(unsigned));
        const unsigned bt = (unsigned)(nk * sizeof(int));
        const unsigned bk = (unsigned)(nk * sizeof(CarryT));
        const unsigned bar = full0 + 8 * s;
        // few knots in the tile and a knot_ls_kernel table for this signal: {L_k, slope_k} of knots kb .. kb + cnt + 1
        // arrive with the tile and no warp evaluates a knot baseline (the tau slice is then not needed)
        const bool lsm = ls_ok && cnt <= kLsTile;
        sm.stage[s].ls_mode = lsm ? 1 : 0;
        if (lsm) {
            const int lq = (sizeof(LS) == 8) ? (kb & ~1) : kb;       // 16-byte aligned start (float pairs are 8 bytes)
            const int ne = min(kb + cnt + 1, K + 1) - lq + 1;
            const unsigned bl = (unsigned)(((sizeof(LS) == 8) ? ((ne + 1) & ~1) : ne) * sizeof(LS));
            mbar_arrive_expect_tx(bar, bx + bm + bk + bl);
            tma_load_1d(st + (unsigned)offsetof(typename Smem::Stage, lsg), gls + lq, bl, bar);
        } else {
            mbar_arrive_expect_tx(bar, bx + bm + bt + bk);
            tma_load_1d(st + (unsigned)offsetof(typename Smem::Stage, tau), gtau + lo, bt, bar);
        }
        tma_load_1d(st + (unsigned)offsetof(typename Smem::Stage, x), x + t0, bx, bar);
        tma_load_1d(st + (unsigned)offsetof(typename Smem::Stage, mask), gmask_in + (t0 >> 5), bm, bar);
        tma_load_1d(st + (unsigned)offsetof(typename Smem::Stage, xk), gxk + lo, bk, bar);
    };
    if (tid == 0)
        for (int i = 0; i < STAGES && i < tiles; ++i) issue_tile(i);

    const int span0 = warp * SPAN;                        // first sample of this warp's span (in tile)
    // running per-thread output pointers (advance by T per tile)
    OutT *rot = reinterpret_cast<OutT *>(p.rot) + row_off + (long long)e * n + span0 + lane;
    OutT *bas = BAS ? reinterpret_cast<OutT *>(p.bas) + row_off + (long long)e * n + span0 + lane : nullptr;
    CarryT *carry = reinterpret_cast<CarryT *>(p.carry_out) + (long long)sig * n + span0 + lane;
    int *ntau = p.next.tau + (long long)sig * p.next.kstride;
    CarryT *nxk = reinterpret_cast<CarryT *>(p.next.xk) + (long long)sig * p.next.kstride;
    unsigned *nmask = p.next.mask + (long long)sig * p.next.mstride + warp * ITEMS + lane;
    int *ntbase = p.next.tbase + (long long)sig * (tiles + 1);
    CarryT *nendl = reinterpret_cast<CarryT *>(p.next.endl) + 2ll * sig;
    LS *ls = sm.ls[warp];
    const unsigned le_mask = 0xffffffffu >> (31 - lane);

    int run_total = 0;          // new-level knots found in earlier tiles (identical in every warp)
    int cached_wb = -1;         // segment whose (L, slope) sit in ls[0..1] from a knot-free span
    bool zero_dx = false;

    auto tile_body = [&](auto edge_tag, const int i) {
        constexpr bool EDGE = decltype(edge_tag)::value;
        const int s = i % STAGES;
        typename Smem::Stage &st = sm.stage[s];
        const int t0 = i * T;
        const int len = EDGE ? min(T, n - t0) : T;
        const int kb = sm.tbase[i];
        const int lo = max(kb - 1, 0) & ~3;
        mbar_wait(full0 + 8 * s, (i / STAGES) & 1);

        // ---- A. segment bases from the stored flag words --------------------------------------
        const int nwords = (len + 31) >> 5;
        const unsigned word = (!EDGE || lane < nwords) ? st.mask[lane] : 0u;
        // knots of the tile before this warp's span
        const int wpre0 = __reduce_add_sync(0xffffffffu, (lane < warp * ITEMS) ? __popc(word) : 0);
        unsigned mw[ITEMS];
        {
            const uint4 q = *reinterpret_cast<const uint4 *>(&st.mask[warp * ITEMS]);
            mw[0] = q.x; mw[1] = q.y; mw[2] = q.z; mw[3] = q.w;
            if (EDGE) {
#pragma unroll
                for (int r = 0; r < ITEMS; ++r) mw[r] = (warp * ITEMS + r < nwords) ? mw[r] : 0u;
            }
        }
        int wpre[ITEMS];                                              // knots of the span before word r
        wpre[0] = 0;
#pragma unroll
        for (int r = 1; r < ITEMS; ++r) wpre[r] = wpre[r - 1] + __popc(mw[r - 1]);
        const int wb = kb + wpre0;                                    // knots before the span = seg(span0 - 1)
        const int wcnt = wpre[ITEMS - 1] + __popc(mw[ITEMS - 1]);     // knots inside the span
        const bool span_live = !EDGE || span0 < len;

        // right-halo sample (first sample after the span)
        const int tend = t0 + span0 + SPAN;                           // its global index
        const bool have_right = !EDGE || (span_live && tend <= n - 1);
        CarryT xright = (CarryT)0;
        int fright = 0;
        if (have_right) {
            if (warp < WARPS - 1) {
                xright = (CarryT)st.x[span0 + SPAN];
                fright = (int)(st.mask[(warp + 1) * ITEMS] & 1u);
            } else {
                // first sample of the NEXT tile: its load was issued STAGES-1 tiles ago
                const int s2 = (i + 1) % STAGES;
                mbar_wait(full0 + 8 * s2, ((i + 1) / STAGES) & 1);
                xright = (CarryT)sm.stage[s2].x[0];
                fright = (int)(sm.stage[s2].mask[0] & 1u);
            }
        }

        // ---- B. knot baseline + slopes for the knots this span touches (warp-private) ----------
        // L for k in [wb, wb + wcnt + 2], slope for segments [wb, wb + wcnt + 1], clipped to the table.
        // A knot-free span inside the same segment as the previous tile reuses the cached pair.
        const CarryT *xkb = st.xk + (wb - lo);                        // xkb[j] = X of knot wb + j
        const bool knot_free = (wcnt == 0 && fright == 0);
        const bool lsm = st.ls_mode != 0;                             // block-uniform
        const LS *lsp = lsm ? st.lsg + (wb - ((sizeof(LS) == 8) ? (kb & ~1) : kb)) : ls;   // lsp[j] = {L, slope} of knot wb + j
        if (!lsm && span_live && !(knot_free && wb == cached_wb)) {
            const int *taub = st.tau + (wb - lo);
            const int nl = min(wcnt + 3, K + 2 - wb);
            const int ns = min(wcnt + 2, K + 1 - wb);
            auto knot_L = [&](const int j) -> CarryT {                // ITD.py:100-110
                const int k = wb + j;
                if (k == 0) return sm.endl[0];
                if (k == K + 1) return sm.endl[1];
                const CarryT w = A::ratio(taub[j] - taub[j - 1], taub[j + 1] - taub[j - 1]);
                const CarryT d = A::sub(xkb[j + 1], xkb[j - 1]);
                const CarryT qq = A::add(xkb[j - 1], A::mul(w, d));
                return A::add(A::mul((CarryT)0.5, qq), A::mul((CarryT)0.5, xkb[j]));
            };
            if (nl <= 32) {
                // one knot per lane; the neighbour's L comes by shuffle
                const CarryT L = (lane < nl) ? knot_L(lane) : (CarryT)0;
                const CarryT Ln = __shfl_down_sync(0xffffffffu, L, 1);
                CarryT sl = (CarryT)0;
                if (lane < ns) {
                    const CarryT den = A::sub(xkb[lane + 1], xkb[lane]);     // ITD.py:116
                    sl = A::div(A::sub(Ln, L), den);
                    zero_dx |= (den == (CarryT)0);
                }
                if (lane < nl) ls[lane] = LS{L, sl};
            } else {
                for (int j = lane; j < nl; j += 32) ls[j].L = knot_L(j);
                __syncwarp();
                for (int j = lane; j < ns; j += 32) {
                    const CarryT den = A::sub(xkb[j + 1], xkb[j]);
                    ls[j].s = A::div(A::sub(ls[j + 1].L, ls[j].L), den);
                    zero_dx |= (den == (CarryT)0);
                }
            }
            __syncwarp();
        }
        cached_wb = (!lsm && span_live && knot_free) ? wb : -1;

        // ---- C. B, R for the span (+ one halo sample each side) --------------------------------
        CarryT b[ITEMS];
        const InT *xs = st.x + span0 + lane;
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) {
            const int jt = span0 + r * 32 + lane;
            CarryT bv = (CarryT)0;
            if (!EDGE || jt < len) {
                const CarryT xv = (CarryT)xs[r * 32];
                const int j = wpre[r] + __popc(mw[r] & le_mask);
                const LS q = lsp[j];
                bv = A::add(q.L, A::mul(q.s, A::sub(xv, xkb[j])));    // ITD.py:115-117
                if (EDGE && t0 + jt == n - 1) bv = (CarryT)0;         // ITD.py:112
                const CarryT rr = A::sub(xv, bv);
                rot[r * 32] = (OutT)(LAST ? A::add(rr, bv) : rr);     // ITD.py:119 / :420
                carry[r * 32] = bv;
                if (BAS) bas[r * 32] = LAST ? (OutT)0 : (OutT)bv;     // ITD.py:424
                if (EDGE && t0 + jt == n - 2) nendl[1] = mean2<CarryT>(bv, (CarryT)0);
            }
            b[r] = bv;
        }
        rot += T;
        carry += T;
        if (BAS) bas += T;
        // left halo B[span0 - 1]: previous warp's last sample (same tile) or the previous tile's
        CarryT bleft = (CarryT)0;
        if (span_live) {
            if (warp == 0) {
                bleft = sm.carry_b[(i + 1) & 1];
            } else {
                const CarryT xl = (CarryT)st.x[span0 - 1];
                const LS q = lsp[0];
                bleft = A::add(q.L, A::mul(q.s, A::sub(xl, xkb[0])));
            }
        }
        CarryT bright = (CarryT)0;
        if (have_right && (!EDGE || tend < n - 1)) {
            const int j = wcnt + fright;
            const LS q = lsp[j];
            bright = A::add(q.L, A::mul(q.s, A::sub(xright, xkb[j])));
        }

        // ---- D. extrema of B: next level's flag words -----------------------------------------
        unsigned fw[ITEMS];
        const int newc = span_extrema<EDGE, ITEMS, CarryT>(b, bleft, bright, lane, t0 + span0, n, fw);
        if (lane < ITEMS && (!EDGE || span0 + lane * 32 < len)) {
            unsigned v = fw[0];
#pragma unroll
            for (int r = 1; r < ITEMS; ++r) v = (lane == r) ? fw[r] : v;
            nmask[0] = v;
        }
        nmask += T / 32;
        if (warp == WARPS - 1 && lane == 31) sm.carry_b[i & 1] = b[ITEMS - 1];
        if (EDGE && i == 0 && warp == 0) {
            const CarryT b1 = shfl_idx(b[0], 1);
            if (lane == 0) {
                ntau[0] = 0;
                nxk[0] = b[0];
                nendl[0] = mean2<CarryT>(b[0], b1);
            }
        }
        // ---- E/F/G. one block barrier, then compact the new knots ------------------------------
        compact_knots<WARPS, ITEMS, CarryT>(sm.cnt, i, warp, lane, newc, fw, b, t0 + span0, run_total, ntau, nxk,
                                            ntbase);
        // every warp is past its reads of stage s: refill it with the tile STAGES ahead
        if (tid == 0 && i + STAGES < tiles) issue_tile(i + STAGES);
    };

    for (int i = 0; i < tiles; ++i) {
        if (i == 0 || i == tiles - 1)
            tile_body(std::true_type{}, i);
        else
            tile_body(std::false_type{}, i);
    }

    if (zero_dx) atomicOr(p.status + sig, kStZeroDx);
    if (warp == 0 && lane == 0) {
        const int Kn = run_total;
        ntbase[tiles] = Kn;
        p.next.kcount[sig] = Kn;
        ntau[Kn + 1] = n - 1;
        nxk[Kn + 1] = (CarryT)0;                                      // B[n-1] == 0
        p.knot_counts[(long long)sig * p.rows + e] = Kn;              // what ITD.py:403 prints
        if (Kn < p.min_extrema) {                                     // ITD.py:404
            p.stop_kind[sig] = kStopKnots;
            p.n_rows[sig] = e + 1;
            p.stop_e[sig] = e;
        } else if (LAST) {                                            // ITD.py:418
            p.stop_kind[sig] = kStopIter;
            p.n_rows[sig] = e + 1;
            p.stop_e[sig] = e;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// scan_stream_kernel: extrema detection + compaction on the raw input (ITD.py:87-98), same
// pipeline, one CTA per signal.  Writes the same table/mask/tbase/endl as knot_scan_kernel.
// ---------------------------------------------------------------------------------------------
template <typename InT, typename CarryT, int WARPS, int ITEMS, int STAGES>
__global__ void __launch_bounds__(WARPS * 32, 3) scan_stream_kernel(const ScanParams p) {
    using Smem = StreamSmem<InT, CarryT, WARPS, ITEMS, STAGES, false>;
    constexpr int T = Smem::T;
    constexpr int SPAN = Smem::SPAN;
    extern __shared__ __align__(128) unsigned char smem_stream_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_stream_raw);
    unsigned sbase = smem_u32(smem_stream_raw);
    asm volatile("" : "+r"(sbase));
    const unsigned full0 = sbase + (unsigned)offsetof(Smem, full);

    const int sig = blockIdx.x + p.sig0;
    const int n = p.n, tiles = p.tiles;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        sm.carry_b[0] = sm.carry_b[1] = (CarryT)0;
        for (int s = 0; s < STAGES; ++s) mbar_init(full0 + 8 * s, 1);
        mbar_fence_init();
    }
    __syncthreads();
    const InT *x = reinterpret_cast<const InT *>(p.x) + (long long)sig * n;
    auto issue_tile = [&](const int i) {
        const int s = i % STAGES;
        const unsigned st = sbase + (unsigned)(offsetof(Smem, stage) + (size_t)s * sizeof(typename Smem::Stage));
        const int t0 = i * T;
        const unsigned bx = (unsigned)(min(T, n - t0) * sizeof(InT));
        mbar_arrive_expect_tx(full0 + 8 * s, bx);
        tma_load_1d(st + (unsigned)offsetof(typename Smem::Stage, x), x + t0, bx, full0 + 8 * s);
    };
    if (tid == 0)
        for (int i = 0; i < STAGES && i < tiles; ++i) issue_tile(i);

    const int span0 = warp * SPAN;
    int *ntau = p.out.tau + (long long)sig * p.out.kstride;
    CarryT *nxk = reinterpret_cast<CarryT *>(p.out.xk) + (long long)sig * p.out.kstride;
    unsigned *nmask = p.out.mask + (long long)sig * p.out.mstride + warp * ITEMS + lane;
    int *ntbase = p.out.tbase + (long long)sig * (tiles + 1);
    CarryT *nendl = reinterpret_cast<CarryT *>(p.out.endl) + 2ll * sig;
    int run_total = 0;
    bool bad = false;

    auto tile_body = [&](auto edge_tag, const int i) {
        constexpr bool EDGE = decltype(edge_tag)::value;
        const int s = i % STAGES;
        typename Smem::Stage &st = sm.stage[s];
        const int t0 = i * T;
        const int len = EDGE ? min(T, n - t0) : T;
        mbar_wait(full0 + 8 * s, (i / STAGES) & 1);
        const bool span_live = !EDGE || span0 < len;
        const int tend = t0 + span0 + SPAN;
        const bool have_right = !EDGE || (span_live && tend <= n - 1);
        CarryT v[ITEMS];
        const InT *xs = st.x + span0 + lane;
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) {
            const int jt = span0 + r * 32 + lane;
            v[r] = (!EDGE || jt < len) ? (CarryT)xs[r * 32] : (CarryT)0;
            bad |= !isfinite(v[r]);
            if (EDGE && t0 + jt == n - 2) {
                const CarryT xl = (CarryT)st.x[jt + 1 < len ? jt + 1 : jt];   // x[n-1] is in this tile iff jt+1 < len
                if (jt + 1 < len) nendl[1] = mean2<CarryT>(v[r], xl);
            }
        }
        CarryT vleft = (CarryT)0, vright = (CarryT)0;
        if (span_live) vleft = (warp == 0) ? sm.carry_b[(i + 1) & 1] : (CarryT)st.x[span0 - 1];
        if (have_right) {
            if (warp < WARPS - 1) {
                vright = (CarryT)st.x[span0 + SPAN];
            } else {
                const int s2 = (i + 1) % STAGES;
                mbar_wait(full0 + 8 * s2, ((i + 1) / STAGES) & 1);
                vright = (CarryT)sm.stage[s2].x[0];
            }
        }
        // x[n-1] opens a tile of its own: x[n-2] is the previous tile's last sample
        if (EDGE && t0 == n - 1 && warp == 0 && lane == 0) nendl[1] = mean2<CarryT>(vleft, v[0]);

        unsigned fw[ITEMS];
        int newc = span_extrema<EDGE, ITEMS, CarryT>(v, vleft, vright, lane, t0 + span0, n, fw);
        if (p.kinds != 3) {
            // detect_peaks(x) alone (valleys) or detect_peaks(-x) alone (peaks): filter the union
            newc = 0;
#pragma unroll
            for (int r = 0; r < ITEMS; ++r) {
                const CarryT give = (lane == 0) ? ((r + 1 < ITEMS) ? v[(r + 1 < ITEMS) ? r + 1 : r] : vright) : v[r];
                const CarryT nx = shfl_idx(give, (lane + 1) & 31);
                const unsigned LT = __ballot_sync(0xffffffffu, v[r] < nx);   // rising after the sample = valley
                fw[r] &= (p.kinds == 1) ? LT : ~LT;
                newc += __popc(fw[r]);
            }
        }
        if (lane < ITEMS && (!EDGE || span0 + lane * 32 < len)) {
            unsigned w = fw[0];
#pragma unroll
            for (int r = 1; r < ITEMS; ++r) w = (lane == r) ? fw[r] : w;
            nmask[0] = w;
        }
        nmask += T / 32;
        if (warp == WARPS - 1 && lane == 31) sm.carry_b[i & 1] = v[ITEMS - 1];
        if (EDGE && i == 0 && warp == 0) {
            const CarryT v1 = shfl_idx(v[0], 1);
            if (lane == 0) {
                ntau[0] = 0;
                nxk[0] = v[0];
                nendl[0] = mean2<CarryT>(v[0], v1);
            }
        }
        if (EDGE && t0 + span0 <= n - 1 && n - 1 < t0 + span0 + SPAN) {
            // the lane holding x[n-1] publishes the closing knot's value (index written at the end)
            const int q = n - 1 - t0 - span0;
#pragma unroll
            for (int r = 0; r < ITEMS; ++r)
                if (q == r * 32 + lane) sm.endl[0] = v[r];
        }
        compact_knots<WARPS, ITEMS, CarryT>(sm.cnt, i, warp, lane, newc, fw, v, t0 + span0, run_total, ntau, nxk,
                                            ntbase);
        if (tid == 0 && i + STAGES < tiles) issue_tile(i + STAGES);
    };

    for (int i = 0; i < tiles; ++i) {
        if (i == 0 || i == tiles - 1)
            tile_body(std::true_type{}, i);
        else
            tile_body(std::false_type{}, i);
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(p.status + sig, kStNonFinite);
    named_barrier_sync(1, WARPS * 32);                  // sm.endl[0] (x[n-1]) is visible
    if (warp == 0 && lane == 0) {
        const int K = run_total;
        ntbase[tiles] = K;
        p.out.kcount[sig] = K;
        ntau[K + 1] = n - 1;
        nxk[K + 1] = sm.endl[0];
        if (p.input_knots) p.input_knots[sig] = K;
    }
}

}  // namespace pyitd
