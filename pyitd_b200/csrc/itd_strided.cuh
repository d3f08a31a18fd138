// itd_strided.cuh -- ONE LONG SIGNAL (BASELINE config 3: 2^28 samples) with the streaming tile pipeline.
//
// level_stream_kernel gives one CTA a whole signal and walks its tiles in order; a single long signal would
// occupy one CTA.  The multi-CTA look-back kernel (itd_kernels.cuh) spreads the tiles over the GPU but runs
// one tile per CTA: every CTA pays its launch, an un-prefetched load of the tile, eight block barriers and
// exits (measured 3.7-4.3 ms per level of 2^28 samples, 0.19 of the HBM roofline).
//
// level_strided_kernel is the streaming kernel made persistent over ONE signal: G = (CTAs that fit on the
// GPU) blocks, block c takes tiles c, c + G, c + 2G, ... in increasing order.  Per tile it is the same
// pipeline as level_stream_kernel -- TMA ring for the samples, the flag words and the knot-table slice,
// warp-private knot baseline / slopes (ITD.py:106-110, :116), B, R, the stencil on B -- with two changes
// that remove the in-order carry between neighbouring tiles:
//   * the level kernel cannot know the global rank of a new knot: that needs the number of knots in ALL earlier tiles,
//     and a look-back chain inside a persistent grid puts every block in lock-step with the slowest one (measured:
//     5.1 ms per level, worse than one CTA per tile).  So it compacts the new knots INSIDE the tile -- (tau, X) go to
//     the tile's own slot [i*T, i*T + count_i) of the staging arrays, values straight from registers -- stores the flag
//     words and the tile's count, and adds the count to the sum of its group of 32 tiles (one reduction per non-empty
//     tile).  tile_prefix_kernel scans the group sums (8192 for 2^28 samples) and applies the stop rule
//     (ITD.py:400-404 / :418); place_knots_kernel finishes the per-tile prefix and moves every slot to its global rank
//     with contiguous loads and stores: the second half of the "extrema-compaction pass";
//   * the halo samples x[t0 - 1], x[t0 + T] and the flag of sample t0 + T ride in with the tile (the TMA
//     slice is four samples / four flag words wider on each side), so the first warp evaluates B[t0 - 1]
//     from the knot table like every other warp and nobody peeks into a neighbour's stage.
// Arithmetic, operation order and every stored table are those of level_stream_kernel: results are
// bit-identical (tests/test_gpu_parity.py::test_strided_long_signal_kernel).
#pragma once

#include "itd_stream.cuh"

namespace pyitd {

template <typename InT, typename CarryT, int WARPS, int ITEMS, int STAGES>
struct StridedSmem {
    static constexpr int T = WARPS * 32 * ITEMS;
    static constexpr int KC = T + 16;                   // knot slice capacity (T + 5, start aligned down to 4)
    static constexpr int SPAN = 32 * ITEMS;
    static constexpr int SC = SPAN + 8;                 // per-warp knot scratch
    struct alignas(2 * sizeof(CarryT)) LS {
        CarryT L, s;
    };
    struct Stage {
        alignas(16) InT xpad[4];                        // x[t0 - 4 .. t0 - 1]
        InT x[T + 4];                                   // x[t0 .. t0 + T + 3]
        alignas(16) unsigned mask[T / 32 + 4];          // flag words of the tile + the next tile's first word
        alignas(16) int tau[KC];
        alignas(16) CarryT xk[KC];
        alignas(16) LS lsg[kLsStage];                   // slice of the knot_ls_kernel table (tiles with few knots)
        int kb, cnt;                                    // knots before / inside the tile (from the per-tile prefix)
        int ls_mode;                                    // 1: lsg holds {L, slope} of knots kb .. kb + cnt + 1
    };
    Stage stage[STAGES];
    alignas(8) unsigned long long full[STAGES];
    LS ls[WARPS][SC];
    int cnt[2][WARPS];
    CarryT endl[2];
};
static_assert(sizeof(float) == 4 && sizeof(double) == 8, "");

template <typename InT, typename CarryT, typename OutT, int WARPS, int ITEMS, int STAGES, bool LAST, bool BAS>
__global__ void __launch_bounds__(WARPS * 32, 3) level_strided_kernel(const LevelParams p) {
    using A = Arith<CarryT>;
    using Smem = StridedSmem<InT, CarryT, WARPS, ITEMS, STAGES>;
    using LS = typename Smem::LS;
    using Stage = typename Smem::Stage;
    constexpr int T = Smem::T;
    constexpr int SPAN = Smem::SPAN;
    static_assert(T / 32 <= 32 && ITEMS == 4, "one flag word per lane, four words per warp (LDS.128)");
    extern __shared__ __align__(128) unsigned char smem_strided_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_strided_raw);
    unsigned sbase = smem_u32(smem_strided_raw);
    asm volatile("" : "+r"(sbase));
    const unsigned full0 = sbase + (unsigned)offsetof(Smem, full);

    const int sig = p.sig0;                               // the one signal of this launch
    const int n = p.n, e = p.e, tiles = p.tiles;
    const int G = (int)gridDim.x, cta = (int)blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // tile k of this block is tile first + k * tstride of the signal: every G-th tile (neighbouring blocks stream
    // neighbouring tiles), or a contiguous run (kOptContigTiles: the knot-free-span cache of section B then survives
    // from one tile to the next, as in level_stream_kernel)
    const bool contig = (p.opts & kOptContigTiles) != 0;
    const int per = (tiles + G - 1) / G;
    const int first = contig ? cta * per : cta;
    const int tstride = contig ? 1 : G;
    const int my_tiles = contig ? max(0, min(per, tiles - first)) : (tiles - cta + G - 1) / G;
    const long long row_off = (long long)sig * p.out_sig_stride;

    // ---- a signal that already stopped: trend-row fix-up / zero tail, tile by tile ----------------
    const int se = p.stop_e[sig];
    if (e > se) {
        OutT *rot = reinterpret_cast<OutT *>(p.rot) + row_off;
        OutT *bas = BAS ? reinterpret_cast<OutT *>(p.bas) + row_off : nullptr;
        const CarryT *src = reinterpret_cast<const CarryT *>(p.fix_src) + (long long)sig * n;
        const bool fix = (e == se + 1 && p.stop_kind[sig] == kStopKnots);
        const bool ztail = (p.opts & kOptZeroTail) && e < p.rows;
        if (fix || ztail) {
            for (int i = cta; i < tiles; i += G) {
                const int t0 = i * T, len = min(T, n - t0);
                if (fix) {
                    // the discarded extraction `se` wrote R_se into row se; the reference returns
                    // baselines[se-1] there (ITD.py:410-411), i.e. the INPUT of that extraction (zeros when se == 0)
                    copy_row(rot + (long long)se * n + t0, src + t0, len, se == 0);
                    if (BAS && (p.opts & kOptZeroTail)) copy_row(bas + (long long)se * n + t0, src + t0, len, true);
                }
                if (ztail) {
                    copy_row(rot + (long long)e * n + t0, src + t0, len, true);
                    if (BAS) copy_row(bas + (long long)e * n + t0, src + t0, len, true);
                }
            }
        }
        return;
    }
    if (e > p.emax) return;

    // ---- prologue ---------------------------------------------------------------------------
    const int K = p.cur.kcount[sig];
    if (tid == 0) {
        const CarryT *gendl = reinterpret_cast<const CarryT *>(p.cur.endl) + 2ll * sig;
        sm.endl[0] = gendl[0];
        sm.endl[1] = gendl[1];
        for (int s = 0; s < STAGES; ++s) mbar_init(full0 + 8 * s, 1);
        mbar_fence_init();
    }
    __syncthreads();

    const InT *x = reinterpret_cast<const InT *>(p.in) + (long long)sig * n;
    const int *gtb = p.cur.tbase + (long long)sig * (tiles + 1);
    const int *gtau = p.cur.tau + (long long)sig * p.cur.kstride;
    const CarryT *gxk = reinterpret_cast<const CarryT *>(p.cur.xk) + (long long)sig * p.cur.kstride;
    const unsigned *gmask_in = p.cur.mask + (long long)sig * p.cur.mstride;
    const LS *gls = reinterpret_cast<const LS *>(p.ls) + (long long)sig * (p.lscap + 4);
    const bool ls_ok = (p.ls != nullptr) && K <= p.lscap;
    // TMA bulk loads of this block's k-th tile into stage k % STAGES (one thread)
    auto issue_tile = [&](const int k, const int kb, const int kb1) {
        const int i = first + k * tstride;
        const int s = k % STAGES;
        Stage &sg = sm.stage[s];
        const unsigned st = sbase + (unsigned)(offsetof(Smem, stage) + (size_t)s * sizeof(Stage));
        const int t0 = i * T;
        const int len = min(T, n - t0);
        const int cnt = kb1 - kb;
        sg.kb = kb;
        sg.cnt = cnt;
        const int lo = max(kb - 1, 0) & ~3;
        const int hi = min(kb + cnt + 3, K + 1);
        const int nk = (hi - lo + 1 + 3) & ~3;
        const int left = (i > 0) ? 4 : 0;                               // halo samples before the tile
        const int right = (t0 + T < n) ? 4 : 0;                         // ... and after it
        const unsigned bx = (unsigned)((left + len + right) * sizeof(InT));
        const unsigned bm = (unsigned)(((((len + 31) / 32 + 3) & ~3) + right) * sizeof(unsigned));
        const unsigned bt = (unsigned)(nk * sizeof(int));
        const unsigned bk = (unsigned)(nk * sizeof(CarryT));
        const unsigned bar = full0 + 8 * s;
        // few knots in the tile and a knot_ls_kernel table: {L_k, slope_k} of knots kb .. kb + cnt + 1 arrive with the tile
        const bool lsm = ls_ok && cnt <= kLsTile;
        sg.ls_mode = lsm ? 1 : 0;
        if (lsm) {
            const int lq = (sizeof(LS) == 8) ? (kb & ~1) : kb;       // 16-byte aligned start (float pairs are 8 bytes)
            const int ne = min(kb + cnt + 1, K + 1) - lq + 1;
            const unsigned bl = (unsigned)(((sizeof(LS) == 8) ? ((ne + 1) & ~1) : ne) * sizeof(LS));
            mbar_arrive_expect_tx(bar, bx + bm + bk + bl);
            tma_load_1d(st + (unsigned)offsetof(Stage, lsg), gls + lq, bl, bar);
        } else {
            mbar_arrive_expect_tx(bar, bx + bm + bt + bk);
            tma_load_1d(st + (unsigned)offsetof(Stage, tau), gtau + lo, bt, bar);
        }
        tma_load_1d(st + (unsigned)(offsetof(Stage, x) - left * sizeof(InT)), x + t0 - left, bx, bar);
        tma_load_1d(st + (unsigned)offsetof(Stage, mask), gmask_in + (t0 >> 5), bm, bar);
        tma_load_1d(st + (unsigned)offsetof(Stage, xk), gxk + lo, bk, bar);
    };
    if (tid == 0)
        for (int k = 0; k < STAGES && k < my_tiles; ++k)
            issue_tile(k, gtb[first + k * tstride], gtb[first + k * tstride + 1]);
    int pf_kb = 0, pf_kb1 = 0;     // thread 0: per-tile knot prefix of the tile it will issue after this one (loaded early)

    const int span0 = warp * SPAN;                        // first sample of this warp's span (in tile)
    OutT *rot0 = reinterpret_cast<OutT *>(p.rot) + row_off + (long long)e * n + span0 + lane;
    OutT *bas0 = BAS ? reinterpret_cast<OutT *>(p.bas) + row_off + (long long)e * n + span0 + lane : nullptr;
    CarryT *carry0 = reinterpret_cast<CarryT *>(p.carry_out) + (long long)sig * n + span0 + lane;
    int *ntau = p.next.tau + (long long)sig * p.next.kstride;
    CarryT *nxk = reinterpret_cast<CarryT *>(p.next.xk) + (long long)sig * p.next.kstride;
    unsigned *nmask0 = p.next.mask + (long long)sig * p.next.mstride + warp * ITEMS + lane;
    int *ntbase = p.next.tbase + (long long)sig * (tiles + 1);
    int *ngsum = p.next.gsum + (long long)sig * p.next.gstride;
    int *stau = p.next.stau + (long long)sig * p.next.kstride;
    CarryT *sxk = reinterpret_cast<CarryT *>(p.next.sxk) + (long long)sig * p.next.kstride;
    CarryT *nendl = reinterpret_cast<CarryT *>(p.next.endl) + 2ll * sig;
    LS *ls = sm.ls[warp];
    const unsigned le_mask = 0xffffffffu >> (31 - lane);

    int cached_wb = -1;         // segment whose (L, slope) sit in ls[0..1] from a knot-free span
    bool zero_dx = false;

    auto tile_body = [&](auto edge_tag, const int k) {
        constexpr bool EDGE = decltype(edge_tag)::value;
        const int i = first + k * tstride;
        const int s = k % STAGES;
        Stage &st = sm.stage[s];
        const int t0 = i * T;
        const int len = EDGE ? min(T, n - t0) : T;
        mbar_wait(full0 + 8 * s, (k / STAGES) & 1);
        const InT *xt = st.xpad + 4;                                  // xt[j] = x[t0 + j], j in [-4, T + 4)
        const int kb = st.kb;
        const int lo = max(kb - 1, 0) & ~3;
        OutT *rot = rot0 + t0;
        OutT *bas = BAS ? bas0 + t0 : nullptr;
        CarryT *carry = carry0 + t0;

        // ---- A. segment bases from the stored flag words --------------------------------------
        const int nwords = (len + 31) >> 5;
        const unsigned word = (!EDGE || lane < nwords) ? st.mask[lane] : 0u;
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

        // right-halo sample (first sample after the span): inside the tile, or the halo that came with it
        const int tend = t0 + span0 + SPAN;                           // its global index
        const bool have_right = !EDGE || (span_live && tend <= n - 1);
        CarryT xright = (CarryT)0;
        int fright = 0;
        if (have_right) {
            xright = (CarryT)xt[span0 + SPAN];
            fright = (int)(st.mask[(warp + 1) * ITEMS] & 1u);
        }

        // ---- B. knot baseline + slopes for the knots this span touches (warp-private) ----------
        const CarryT *xkb = st.xk + (wb - lo);                        // xkb[j] = X of knot wb + j
        const bool knot_free = (wcnt == 0 && fright == 0);
        const bool lsm = st.ls_mode != 0;                             // block-uniform
        const LS *lsp = lsm ? st.lsg + (wb - ((sizeof(LS) == 8) ? (kb & ~1) : kb)) : ls;   // lsp[j] = {L, slope} of knot wb + j
        if (!lsm && span_live && !(knot_free && wb == cached_wb)) {
            const int *taub = st.tau + (wb - lo);
            const int nl = min(wcnt + 3, K + 2 - wb);
            const int ns = min(wcnt + 2, K + 1 - wb);
            auto knot_L = [&](const int j) -> CarryT {                // ITD.py:100-110
                const int kk = wb + j;
                if (kk == 0) return sm.endl[0];
                if (kk == K + 1) return sm.endl[1];
                const CarryT w = A::ratio(taub[j] - taub[j - 1], taub[j + 1] - taub[j - 1]);
                const CarryT d = A::sub(xkb[j + 1], xkb[j - 1]);
                const CarryT qq = A::add(xkb[j - 1], A::mul(w, d));
                return A::add(A::mul((CarryT)0.5, qq), A::mul((CarryT)0.5, xkb[j]));
            };
            if (nl <= 32) {
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
        const InT *xs = xt + span0 + lane;
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
        // left halo B[t0 + span0 - 1]: the sample before the span lies in the segment of knot wb (ls[0]);
        // for the first warp it is the halo sample that came in front of the tile
        CarryT bleft = (CarryT)0;
        if (span_live && (t0 + span0 > 0)) {
            const CarryT xl = (CarryT)xt[span0 - 1];
            const LS q = lsp[0];
            bleft = A::add(q.L, A::mul(q.s, A::sub(xl, xkb[0])));
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
            nmask0[t0 >> 5] = v;
        }
        if (EDGE && i == 0 && warp == 0) {
            const CarryT b1 = shfl_idx(b[0], 1);
            if (lane == 0) {
                ntau[0] = 0;
                nxk[0] = b[0];
                nendl[0] = mean2<CarryT>(b[0], b1);
            }
        }
        // ---- E. one block barrier: exchange the per-warp counts, then compact the new knots INSIDE the tile's own
        // slot of the staging arrays (the global ranks are assigned by place_knots_kernel) -------------------------
        if (lane == 0) sm.cnt[k & 1][warp] = newc;
        named_barrier_sync(1, WARPS * 32);               // also: every warp is past its reads of stage s
        if (tid == 0 && k + STAGES < my_tiles) issue_tile(k + STAGES, pf_kb, pf_kb1);
        const int c = (lane < WARPS) ? sm.cnt[k & 1][lane] : 0;
        const int tot = __reduce_add_sync(0xffffffffu, c);
        int pre = t0 + __reduce_add_sync(0xffffffffu, (lane < warp) ? c : 0);
        if (tid == 0) {
            ntbase[i] = tot;                             // place_knots_kernel turns counts into the exclusive prefix
            if (tot) atomicAdd(ngsum + (i >> 5), tot);   // knots per group of 32 tiles (tile_prefix_kernel scans these)
        }
        // (skipping this for knot-free tiles was measured: deep levels -0.02 ms, mid levels +0.05 ms each; not kept)
        const unsigned lt_mask = le_mask >> 1;
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) {
            if ((fw[r] >> lane) & 1u) {
                const int rank = pre + __popc(fw[r] & lt_mask);
                stau[rank] = t0 + span0 + r * 32 + lane;
                sxk[rank] = b[r];
            }
            pre += __popc(fw[r]);
        }
    };

    for (int k = 0; k < my_tiles; ++k) {
        const int i = first + k * tstride;
        if (tid == 0 && k + STAGES < my_tiles) {
            pf_kb = gtb[i + STAGES * tstride];
            pf_kb1 = gtb[i + STAGES * tstride + 1];
        }
        if (i == 0 || i == tiles - 1)
            tile_body(std::true_type{}, k);
        else
            tile_body(std::false_type{}, k);
    }
    if (zero_dx) atomicOr(p.status + sig, kStZeroDx);
}

// ---------------------------------------------------------------------------------------------
// scan_strided_kernel: extrema of the raw input (ITD.py:87-98) for ONE long signal -- flag words and per-tile
// counts + tile-local knot slots (tile_prefix_scan_kernel / place_knots_kernel finish the table).  Same persistent striding
// and halo-carrying TMA slices as level_strided_kernel.
// ---------------------------------------------------------------------------------------------
template <typename InT, int WARPS, int ITEMS>
struct ScanStridedSmem {
    static constexpr int T = WARPS * 32 * ITEMS;
    struct Stage {
        alignas(16) InT xpad[4];
        InT x[T + 4];
    };
    Stage stage[2];
    alignas(8) unsigned long long full[2];
    int cnt[2][WARPS];
};

template <typename InT, typename CarryT, int WARPS, int ITEMS>
__global__ void __launch_bounds__(WARPS * 32, 4) scan_strided_kernel(const ScanParams p) {
    using Smem = ScanStridedSmem<InT, WARPS, ITEMS>;
    using Stage = typename Smem::Stage;
    constexpr int T = Smem::T, SPAN = 32 * ITEMS, STAGES = 2;
    extern __shared__ __align__(128) unsigned char smem_strided_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_strided_raw);
    unsigned sbase = smem_u32(smem_strided_raw);
    asm volatile("" : "+r"(sbase));
    const unsigned full0 = sbase + (unsigned)offsetof(Smem, full);
    const int sig = p.sig0;
    const int n = p.n, tiles = p.tiles;
    const int G = (int)gridDim.x, cta = (int)blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(full0 + 8 * s, 1);
        mbar_fence_init();
    }
    __syncthreads();
    const InT *x = reinterpret_cast<const InT *>(p.x) + (long long)sig * n;
    auto issue_tile = [&](const int k) {
        const int i = cta + k * G, s = k % STAGES;
        const unsigned st = sbase + (unsigned)(offsetof(Smem, stage) + (size_t)s * sizeof(Stage));
        const int t0 = i * T, len = min(T, n - t0);
        const int left = (i > 0) ? 4 : 0, right = (t0 + T < n) ? 4 : 0;
        const unsigned bx = (unsigned)((left + len + right) * sizeof(InT));
        mbar_arrive_expect_tx(full0 + 8 * s, bx);
        tma_load_1d(st + (unsigned)(offsetof(Stage, x) - left * sizeof(InT)), x + t0 - left, bx, full0 + 8 * s);
    };
    const int my_tiles = (tiles - cta + G - 1) / G;
    if (tid == 0)
        for (int k = 0; k < STAGES && k < my_tiles; ++k) issue_tile(k);
    const int span0 = warp * SPAN;
    int *ntau = p.out.tau + (long long)sig * p.out.kstride;
    CarryT *nxk = reinterpret_cast<CarryT *>(p.out.xk) + (long long)sig * p.out.kstride;
    unsigned *nmask0 = p.out.mask + (long long)sig * p.out.mstride + warp * ITEMS + lane;
    int *ntbase = p.out.tbase + (long long)sig * (tiles + 1);
    int *ngsum = p.out.gsum + (long long)sig * p.out.gstride;
    int *stau = p.out.stau + (long long)sig * p.out.kstride;
    CarryT *sxk = reinterpret_cast<CarryT *>(p.out.sxk) + (long long)sig * p.out.kstride;
    CarryT *nendl = reinterpret_cast<CarryT *>(p.out.endl) + 2ll * sig;
    bool bad = false;

    auto tile_body = [&](auto edge_tag, const int k) {
        constexpr bool EDGE = decltype(edge_tag)::value;
        const int i = cta + k * G, s = k % STAGES;
        const int t0 = i * T;
        const int len = EDGE ? min(T, n - t0) : T;
        mbar_wait(full0 + 8 * s, (k / STAGES) & 1);
        const InT *xt = sm.stage[s].xpad + 4;                          // xt[j] = x[t0 + j], j in [-4, T + 4)
        const bool span_live = !EDGE || span0 < len;
        const int tend = t0 + span0 + SPAN;
        const bool have_right = !EDGE || (span_live && tend <= n - 1);
        CarryT v[ITEMS];
        const InT *xs = xt + span0 + lane;
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) {
            const int jt = span0 + r * 32 + lane;
            v[r] = (!EDGE || jt < len) ? (CarryT)xs[r * 32] : (CarryT)0;
            bad |= !isfinite(v[r]);
            if (EDGE && t0 + jt == n - 2 && jt + 1 < len) nendl[1] = mean2<CarryT>(v[r], (CarryT)xt[jt + 1]);   // ITD.py:102
        }
        CarryT vleft = (CarryT)0, vright = (CarryT)0;
        if (span_live && t0 + span0 > 0) vleft = (CarryT)xt[span0 - 1];
        if (have_right) vright = (CarryT)xt[span0 + SPAN];
        // x[n-1] opens a tile of its own: x[n-2] is the halo sample in front of it
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
            nmask0[t0 >> 5] = w;
        }
        if (EDGE && i == 0 && warp == 0) {
            const CarryT v1 = shfl_idx(v[0], 1);
            if (lane == 0) {
                ntau[0] = 0;
                nxk[0] = v[0];
                nendl[0] = mean2<CarryT>(v[0], v1);                      // ITD.py:101
            }
        }
        if (lane == 0) sm.cnt[k & 1][warp] = newc;
        named_barrier_sync(1, WARPS * 32);               // also: every warp is past its reads of stage s
        if (tid == 0 && k + STAGES < my_tiles) issue_tile(k + STAGES);
        const int c = (lane < WARPS) ? sm.cnt[k & 1][lane] : 0;
        const int tot = __reduce_add_sync(0xffffffffu, c);
        int pre = t0 + __reduce_add_sync(0xffffffffu, (lane < warp) ? c : 0);
        if (tid == 0) {
            ntbase[i] = tot;
            if (tot) atomicAdd(ngsum + (i >> 5), tot);
        }
        const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) {                // tile-local compaction (see level_strided_kernel)
            if ((fw[r] >> lane) & 1u) {
                const int rank = pre + __popc(fw[r] & lt_mask);
                stau[rank] = t0 + span0 + r * 32 + lane;
                sxk[rank] = v[r];
            }
            pre += __popc(fw[r]);
        }
    };
    for (int k = 0; k < my_tiles; ++k) {
        const int i = cta + k * G;
        if (i == 0 || i == tiles - 1)
            tile_body(std::true_type{}, k);
        else
            tile_body(std::false_type{}, k);
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(p.status + sig, kStNonFinite);
}

// ---------------------------------------------------------------------------------------------
// Exclusive prefix of the per-GROUP knot sums gsum[0 .. groups) into gbase by ONE block of 1024 threads; gsum is
// cleared on the way (it is accumulated again by the next level kernel); returns the total (valid in every thread).
// A group is 32 tiles, so 2^28 samples are 8192 groups: eight entries per thread.  Warp w owns a contiguous run; a
// lane takes four consecutive entries per round.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int block_group_prefix_1024(int *gsum, int *gbase, const int groups) {
    __shared__ int s_warp[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rounds = (groups + 32 * 128 - 1) / (32 * 128);          // rounds of 128 entries per warp
    const int w0 = warp * rounds * 128;                                 // first entry of this warp's run
    auto load4 = [&](const int r, int (&v)[4]) {
        const int i0 = w0 + r * 128 + lane * 4;
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = (i0 + u < groups) ? gsum[i0 + u] : 0;
    };
    int sum = 0;
    for (int r = 0; r < rounds; ++r) {
        int v[4];
        load4(r, v);
        sum += v[0] + v[1] + v[2] + v[3];
    }
    sum = __reduce_add_sync(0xffffffffu, sum);
    if (lane == 0) s_warp[warp] = sum;
    __syncthreads();
    const int wt = s_warp[lane];
    const int total = __reduce_add_sync(0xffffffffu, wt);
    int run = __reduce_add_sync(0xffffffffu, (lane < warp) ? wt : 0);   // knots before this warp's run
    for (int r = 0; r < rounds; ++r) {
        int v[4];
        load4(r, v);
        const int mine = v[0] + v[1] + v[2] + v[3];
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        int ex = run + incl - mine;
        const int i0 = w0 + r * 128 + lane * 4;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (i0 + u < groups) {
                gbase[i0 + u] = ex;
                gsum[i0 + u] = 0;
            }
            ex += v[u];
        }
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
    return total;
}

// tile_prefix for the scan pass: no stop rule; the closing knot carries x[n-1]; K is the input's knot count
template <typename InT, typename CarryT>
__global__ void __launch_bounds__(1024) tile_prefix_scan_kernel(KnotTable out, const void *xin, int sig, int tiles, int n,
                                                                int *input_knots) {
    const int tid = threadIdx.x;
    int *tb = out.tbase + (long long)sig * (tiles + 1);
    const int K = block_group_prefix_1024(out.gsum + (long long)sig * out.gstride, out.gbase + (long long)sig * out.gstride,
                                          (tiles + 31) >> 5);
    if (tid == 0) {
        const InT *x = reinterpret_cast<const InT *>(xin) + (long long)sig * n;
        int *tau = out.tau + (long long)sig * out.kstride;
        CarryT *xk = reinterpret_cast<CarryT *>(out.xk) + (long long)sig * out.kstride;
        tb[tiles] = K;
        out.kcount[sig] = K;
        tau[K + 1] = n - 1;                                           // ITD.py:98
        xk[K + 1] = (CarryT)x[n - 1];
        if (input_knots) input_knots[sig] = K;
    }
}

// ---------------------------------------------------------------------------------------------
// tile_prefix_kernel: per-group knot sums -> exclusive prefix over the groups, K, the closing knot (ITD.py:98),
// what ITD.py:403 prints and the stop rule (ITD.py:404, :418).  One block.  (The per-tile prefix inside a group is
// finished by the warp of place_knots_kernel that owns the group.)
// ---------------------------------------------------------------------------------------------
template <typename CarryT>
__global__ void __launch_bounds__(1024) tile_prefix_kernel(KnotTable next, int sig, int tiles, int n, int e, int rows,
                                                           int min_extrema, int last, int *stop_e, int *stop_kind,
                                                           int *n_rows, int *knot_counts) {
    const int tid = threadIdx.x;
    if (e > stop_e[sig]) return;                          // the level kernel did nothing either
    int *tb = next.tbase + (long long)sig * (tiles + 1);
    const int Kn = block_group_prefix_1024(next.gsum + (long long)sig * next.gstride,
                                           next.gbase + (long long)sig * next.gstride, (tiles + 31) >> 5);
    if (tid == 0) {
        int *ntau = next.tau + (long long)sig * next.kstride;
        CarryT *nxk = reinterpret_cast<CarryT *>(next.xk) + (long long)sig * next.kstride;
        tb[tiles] = Kn;
        next.kcount[sig] = Kn;
        ntau[Kn + 1] = n - 1;
        nxk[Kn + 1] = (CarryT)0;                                      // B[n-1] == 0 (ITD.py:112)
        knot_counts[(long long)sig * rows + e] = Kn;                  // what ITD.py:403 prints
        if (Kn < min_extrema) {                                       // ITD.py:404
            stop_kind[sig] = kStopKnots;
            n_rows[sig] = e + 1;
            stop_e[sig] = e;
        } else if (last) {                                            // ITD.py:418
            stop_kind[sig] = kStopIter;
            n_rows[sig] = e + 1;
            stop_e[sig] = e;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// place_knots_kernel: the second half of the extrema-compaction pass.  The level / scan kernel left the knots of tile i
// compacted inside the tile's own slot of the staging arrays, stau / sxk [i * T, i * T + count_i); this pass finishes the
// per-tile exclusive prefix (group base + a shuffle scan of 32 counts, stored for the next level kernel) and moves every
// slot to its global rank: tau[1 + rank], xk[1 + rank].  A WARP owns a group of 32 consecutive tiles and walks the
// group's knots 32 at a time -- lane q finds its tile by a five-step search over the 32 prefix values held in the
// warp's registers -- so loads and stores are contiguous runs and nothing waits on a block barrier.
// e_guard: the level whose output is placed; nothing to do once the signal stopped before it.
// ---------------------------------------------------------------------------------------------
template <typename CarryT>
__global__ void __launch_bounds__(256) place_knots_kernel(KnotTable next, int sig, int tiles, int e_guard, const int *stop_e) {
    constexpr int T = 1024;
    if (stop_e && stop_e[sig] < e_guard) return;            // the level kernel of e_guard never ran (stop_e null: scan pass)
    const int lane = threadIdx.x & 31;
    const int gwarp = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), nwarps = (int)((gridDim.x * blockDim.x) >> 5);
    int *tb = next.tbase + (long long)sig * (tiles + 1);
    const int *__restrict__ gbase = next.gbase + (long long)sig * next.gstride;
    const int *__restrict__ stau = next.stau + (long long)sig * next.kstride;
    const CarryT *__restrict__ sxk = reinterpret_cast<const CarryT *>(next.sxk) + (long long)sig * next.kstride;
    int *__restrict__ tau = next.tau + (long long)sig * next.kstride + 1;
    CarryT *__restrict__ xk = reinterpret_cast<CarryT *>(next.xk) + (long long)sig * next.kstride + 1;
    const int groups = (tiles + 31) >> 5;
    for (int g = gwarp; g < groups; g += nwarps) {
        const int i_l = g * 32 + lane;
        const int cnt_l = (i_l < tiles) ? tb[i_l] : 0;
        int incl = cnt_l;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const int ex_l = incl - cnt_l;                                // knots of the group before tile `lane`
        const int base = gbase[g];
        if (i_l < tiles) tb[i_l] = base + ex_l;
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        const long long slot0 = (long long)g * 32 * T;
        for (int q0 = 0; q0 < total; q0 += 128) {
            int src[4], tv[4];
            CarryT xv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int q = q0 + u * 32 + lane;
                int j = 0;                                            // last tile with ex_j <= q (it is never empty)
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) {
                    const int v = __shfl_sync(0xffffffffu, ex_l, (j + step) & 31);
                    if (v <= q) j += step;
                }
                const int ex_j = __shfl_sync(0xffffffffu, ex_l, j);
                src[u] = (q < total) ? j * T + (q - ex_j) : -1;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (src[u] >= 0) {
                    tv[u] = stau[slot0 + src[u]];
                    xv[u] = sxk[slot0 + src[u]];
                }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (src[u] >= 0) {
                    tau[base + q0 + u * 32 + lane] = tv[u];
                    xk[base + q0 + u * 32 + lane] = xv[u];
                }
        }
    }
}

}  // namespace pyitd
