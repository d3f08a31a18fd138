// itd_blk.cuh -- batched-channel level kernel, second generation: ONE CTA PER SIGNAL, BLOCKED lanes,
// and NO knot tables: the only state that travels between levels is the carry X_e and one flag bit per
// sample (the knot mask).
//
// Why (profiles/r1/ncu_stream_v2_metrics.txt, profiles/r2/...): level_stream_kernel is issue-bound at
// 90-160 thread-instructions per sample-level, most of them fixed per-warp-tile work (striped stencil with
// shuffles and ballots, compaction of (tau, X) tables, the count exchange) amortised over only four samples
// per lane.  Here
//   * a lane owns IT CONSECUTIVE samples: the 3-point stencil (ITD.py:44-59) runs in registers with one
//     shuffle per lane instead of one per sample, loads and stores are 128-bit;
//   * a knot is a bit.  Positions come from the mask words, X_k = x[tau_k] is read from the tile that is in
//     shared memory anyway, so the (tau, X) tables, their compaction, the per-tile count exchange and the
//     16-24 B of HBM traffic per knot are gone; the whole-signal mask (N/8 bytes) sits in shared memory;
//   * the two knots before and the two knots after a tile (an unbounded distance away at deep levels) are
//     carried forward by one warp per tile (tile record, one tile ahead); far knots' X are gathered from HBM;
//   * a warp whose span is knot-free and still inside the segment of its previous tile keeps (L_k, slope,
//     X_k) in registers: deep levels run the bare affine map + stencil.
//
// Arithmetic and operation order are those of ITD.py:100-119 (see itd_kernels.cuh); results are bit-identical
// to level_stream_kernel / the oracle.
#pragma once

#include "itd_stream.cuh"

namespace pyitd {

constexpr int kBlkMaxMaskWords = 4096;      // whole-signal mask in shared memory: N <= 131072
constexpr int kNoKnot = 0x7fffffff;         // "no knot after the closing knot"

template <typename CarryT>
struct BlkLS {
    CarryT L;                               // knot baseline L_k (ITD.py:106-110)
    union {
        CarryT s;                           // slope of segment [k, k+1) (ITD.py:116)
        int pos;                            // before the slopes are known: tau_k (global sample index)
    };
};

template <typename CarryT>
struct BlkTileRec {                         // knots around a tile: q[0] nearest before, q[1] the one before it;
    int qpos[2], rpos[2];                   // r[0] first at/after the tile end, r[1] the next one
    CarryT qx[2], rx[2];
};

template <typename InT, typename CarryT, int WARPS, int IT, int STAGES>
struct BlkGeom {
    static constexpr int SPAN = 32 * IT;                 // samples per warp per tile
    static constexpr int T = WARPS * SPAN;               // samples per tile
    static constexpr int NW = T / 32;                    // mask words per tile
    static constexpr int SC = SPAN + 4;                  // knot entries per warp: q1, q0, in-span knots, r0, r1
    static_assert(NW <= 64, "tile mask must fit two words per lane");
    static_assert(IT == 4 || IT == 8, "blocked lanes of 4 or 8 samples");
    struct Misc {
        unsigned long long full[STAGES];
        unsigned long long mfull;
        BlkTileRec<CarryT> tr[2];
        CarryT carry_b[2];
        CarryT endl[2];
        CarryT x0, xlast;
        int cnt[WARPS];
    };
    static constexpr size_t a16(size_t v) { return (v + 15) & ~(size_t)15; }
    static constexpr size_t off_x = 0;
    static constexpr size_t off_ls = a16((size_t)STAGES * T * sizeof(InT));
    static constexpr size_t off_xk = a16(off_ls + (size_t)WARPS * SC * sizeof(BlkLS<CarryT>));
    static constexpr size_t off_misc = a16(off_xk + (size_t)WARPS * SC * sizeof(CarryT));
    static constexpr size_t off_mask = a16(off_misc + sizeof(Misc));
    static size_t bytes(long long mstride) { return off_mask + (size_t)mstride * sizeof(unsigned); }
};

// ---------------------------------------------------------------------------------------------
// knot searches (warp-cooperative, every branch warp-uniform)
// ---------------------------------------------------------------------------------------------
// first knot at a position >= from in the whole-signal mask; the closing knot n-1 (ITD.py:98) when no flag
// is left; kNoKnot past the closing knot
__device__ __forceinline__ int blk_next_knot(const unsigned *mask, int nwords, int from, int n, int lane) {
    if (from > n - 1) return kNoKnot;
    int w0 = from >> 5;
    unsigned keep = 0xffffffffu << (from & 31);
    while (w0 < nwords) {
        const int w = w0 + lane;
        unsigned v = (w < nwords) ? mask[w] : 0u;
        if (lane == 0) v &= keep;
        keep = 0xffffffffu;
        const unsigned bal = __ballot_sync(0xffffffffu, v != 0u);
        if (bal) {
            const int src = __ffs(bal) - 1;
            const unsigned vv = __shfl_sync(0xffffffffu, v, src);
            return ((w0 + src) << 5) + __ffs(vv) - 1;
        }
        w0 += 32;
    }
    return n - 1;
}

struct BlkTileWords {
    unsigned w0, w1;                 // this lane's two mask words of the tile (word lane, word lane + 32)
    unsigned long long nz;           // bit w set <=> word w of the tile has a knot
};
__device__ __forceinline__ unsigned blk_word(const BlkTileWords &tw, int wi) {
    return __shfl_sync(0xffffffffu, (wi & 32) ? tw.w1 : tw.w0, wi & 31);
}
// the two highest flagged samples in words [0, sw) of the tile (tile-relative), -1 when missing
__device__ __forceinline__ void blk_prev2(const BlkTileWords &tw, int sw, int &a, int &b) {
    a = b = -1;
    unsigned long long m = (sw >= 64) ? tw.nz : (tw.nz & ((1ull << sw) - 1ull));
    if (m) {
        int wi = 63 - __clzll((long long)m);
        unsigned v = blk_word(tw, wi);
        const int bit = 31 - __clz((int)v);
        a = wi * 32 + bit;
        v &= ~(1u << bit);
        if (v) {
            b = wi * 32 + 31 - __clz((int)v);
        } else {
            m &= ~(1ull << wi);
            if (m) {
                wi = 63 - __clzll((long long)m);
                v = blk_word(tw, wi);
                b = wi * 32 + 31 - __clz((int)v);
            }
        }
    }
}
// the two lowest flagged samples in words [ew, 64) of the tile
__device__ __forceinline__ void blk_next2(const BlkTileWords &tw, int ew, int &a, int &b) {
    a = b = -1;
    unsigned long long m = (ew >= 64) ? 0ull : (tw.nz & ~((1ull << ew) - 1ull));
    if (m) {
        int wi = __ffsll((long long)m) - 1;
        unsigned v = blk_word(tw, wi);
        a = wi * 32 + __ffs((int)v) - 1;
        v &= v - 1u;
        if (v) {
            b = wi * 32 + __ffs((int)v) - 1;
        } else {
            m &= m - 1ull;
            if (m) {
                wi = __ffsll((long long)m) - 1;
                v = blk_word(tw, wi);
                b = wi * 32 + __ffs((int)v) - 1;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// blocked loads / stores of a lane's IT consecutive samples
// ---------------------------------------------------------------------------------------------
template <int IT, typename CarryT>
__device__ __forceinline__ void blk_load(const double *p, CarryT (&v)[IT]) {
#pragma unroll
    for (int u = 0; u < IT / 2; ++u) {
        const double2 q = reinterpret_cast<const double2 *>(p)[u];
        v[2 * u] = (CarryT)q.x;
        v[2 * u + 1] = (CarryT)q.y;
    }
}
template <int IT, typename CarryT>
__device__ __forceinline__ void blk_load(const float *p, CarryT (&v)[IT]) {
#pragma unroll
    for (int u = 0; u < IT / 4; ++u) {
        const float4 q = reinterpret_cast<const float4 *>(p)[u];
        v[4 * u] = (CarryT)q.x;
        v[4 * u + 1] = (CarryT)q.y;
        v[4 * u + 2] = (CarryT)q.z;
        v[4 * u + 3] = (CarryT)q.w;
    }
}
// nvalid = number of leading samples of the lane that exist (a multiple of 4 because N % 4 == 0)
template <int IT, typename CarryT>
__device__ __forceinline__ void blk_store(double *p, const CarryT (&v)[IT], int nvalid) {
#pragma unroll
    for (int u = 0; u < IT / 2; ++u)
        if (2 * u < nvalid) reinterpret_cast<double2 *>(p)[u] = make_double2((double)v[2 * u], (double)v[2 * u + 1]);
}
template <int IT, typename CarryT>
__device__ __forceinline__ void blk_store(float *p, const CarryT (&v)[IT], int nvalid) {
#pragma unroll
    for (int u = 0; u < IT / 4; ++u)
        if (4 * u < nvalid)
            reinterpret_cast<float4 *>(p)[u] =
                make_float4((float)v[4 * u], (float)v[4 * u + 1], (float)v[4 * u + 2], (float)v[4 * u + 3]);
}

// ---------------------------------------------------------------------------------------------
// level_blk_kernel
// ---------------------------------------------------------------------------------------------
template <typename InT, typename CarryT, typename OutT, int WARPS, int IT, int STAGES, int MINB, bool LAST, bool BAS>
__global__ void __launch_bounds__(WARPS * 32, MINB) level_blk_kernel(const LevelParams p) {
    using A = Arith<CarryT>;
    using G = BlkGeom<InT, CarryT, WARPS, IT, STAGES>;
    using LS = BlkLS<CarryT>;
    using Misc = typename G::Misc;
    constexpr int T = G::T, SPAN = G::SPAN, NW = G::NW, SC = G::SC;
    constexpr unsigned ITMASK = (1u << IT) - 1u;
    extern __shared__ __align__(128) unsigned char smem_blk_raw[];
    unsigned char *sraw = smem_blk_raw;
    InT *sx = reinterpret_cast<InT *>(sraw + G::off_x);
    Misc &mi = *reinterpret_cast<Misc *>(sraw + G::off_misc);
    const unsigned *smask = reinterpret_cast<const unsigned *>(sraw + G::off_mask);
    const unsigned sbase = smem_u32(sraw);
    const unsigned full0 = sbase + (unsigned)(G::off_misc + offsetof(Misc, full));
    const unsigned mfull = sbase + (unsigned)(G::off_misc + offsetof(Misc, mfull));

    const int sig = blockIdx.x + p.sig0;
    const int n = p.n, e = p.e;
    const int tiles = (n + T - 1) / T;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long row_off = (long long)sig * p.out_sig_stride;

    // ---- signals that already stopped (same rules as level_stream_kernel) -----------------------
    const int se = p.stop_e[sig];
    if (e > se) {
        OutT *rot = reinterpret_cast<OutT *>(p.rot) + row_off;
        OutT *bas = BAS ? reinterpret_cast<OutT *>(p.bas) + row_off : nullptr;
        const CarryT *src = reinterpret_cast<const CarryT *>(p.fix_src) + (long long)sig * n;
        if (e == se + 1 && p.stop_kind[sig] == kStopKnots) {
            // the discarded extraction `se` wrote R_se into row se; the reference returns
            // baselines[se-1] there (ITD.py:410-411), i.e. the INPUT of that extraction (zeros when se == 0)
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
    const InT *x = reinterpret_cast<const InT *>(p.in) + (long long)sig * n;
    const int nwords = (int)p.cur.mstride;
    auto issue_tile = [&](const int i) {
        const int s = i % STAGES;
        const int t0 = i * T;
        const unsigned bx = (unsigned)(min(T, n - t0) * sizeof(InT));
        const unsigned bar = full0 + 8 * s;
        mbar_arrive_expect_tx(bar, bx);
        tma_load_1d(sbase + (unsigned)(G::off_x + (size_t)s * T * sizeof(InT)), x + t0, bx, bar);
    };
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(full0 + 8 * s, 1);
        mbar_init(mfull, 1);
        mbar_fence_init();
        const unsigned bm = (unsigned)(nwords * sizeof(unsigned));
        mbar_arrive_expect_tx(mfull, bm);
        tma_load_1d(sbase + (unsigned)G::off_mask, p.cur.mask + (long long)sig * p.cur.mstride, bm, mfull);
        for (int i = 0; i < STAGES && i < tiles; ++i) issue_tile(i);
        mi.carry_b[0] = mi.carry_b[1] = (CarryT)0;
    }
    if (warp == 1 || WARPS == 1) {
        if (lane == 0) {
            const CarryT a0 = (CarryT)x[0], a1 = (CarryT)x[1], z1 = (CarryT)x[n - 2], z0 = (CarryT)x[n - 1];
            mi.endl[0] = mean2<CarryT>(a0, a1);                      // ITD.py:101
            mi.endl[1] = mean2<CarryT>(z1, z0);                      // ITD.py:102
            mi.x0 = a0;
            mi.xlast = z0;
        }
    }
    __syncthreads();                                   // barriers initialised, endl / x0 / xlast visible
    mbar_wait(mfull, 0);                               // the signal's knot mask is in shared memory
    if (warp == 0) {
        // record of tile 0: the opening knot (0, x[0]) before it, the first two knots at/after its end
        const int from = min(T, n - 1);
        const int r0 = blk_next_knot(smask, nwords, from, n, lane);
        const int r1 = (r0 == kNoKnot) ? kNoKnot : blk_next_knot(smask, nwords, r0 + 1, n, lane);
        if (lane == 0) {
            BlkTileRec<CarryT> &tr = mi.tr[0];
            tr.qpos[0] = 0;
            tr.qx[0] = mi.x0;
            tr.qpos[1] = -1;
            tr.qx[1] = (CarryT)0;
            tr.rpos[0] = r0;
            tr.rx[0] = (r0 == kNoKnot) ? (CarryT)0 : (CarryT)x[r0];
            tr.rpos[1] = r1;
            tr.rx[1] = (r1 == kNoKnot) ? (CarryT)0 : (CarryT)x[r1];
        }
    }
    __syncthreads();

    const int span_rel = warp * SPAN;                  // first sample of this warp's span inside a tile
    const int lane_rel = span_rel + lane * IT;         // first sample of this lane inside a tile
    const int sw = warp * IT, ew = sw + IT;            // this span's words inside the tile
    OutT *rot = reinterpret_cast<OutT *>(p.rot) + row_off + (long long)e * n + lane_rel;
    OutT *bas = BAS ? reinterpret_cast<OutT *>(p.bas) + row_off + (long long)e * n + lane_rel : nullptr;
    CarryT *carry = reinterpret_cast<CarryT *>(p.carry_out) + (long long)sig * n + lane_rel;
    unsigned char *nmaskb = reinterpret_cast<unsigned char *>(p.next.mask + (long long)sig * p.next.mstride);
    const int nmask_bits = nwords * 32;                // the mask row covers samples [0, nmask_bits)
    LS *ls = reinterpret_cast<LS *>(sraw + G::off_ls) + warp * SC;
    CarryT *xk = reinterpret_cast<CarryT *>(sraw + G::off_xk) + warp * SC;
    const CarryT endl0 = mi.endl[0], endl1 = mi.endl[1];

    int mycnt = 0;                 // knots of the NEXT level found by this lane
    bool zero_dx = false;
    // segment cache of a knot-free span: (L_k, slope, X_k) of the segment, L of the knot that ends it
    CarryT cL = (CarryT)0, cs = (CarryT)0, cX = (CarryT)0, cLn = (CarryT)0;
    int c_end = -1;                // position of the knot that ends the cached segment (-1: nothing cached)

    auto tile_body = [&](auto edge_tag, const int i) {
        constexpr bool EDGE = decltype(edge_tag)::value;
        const int s = i % STAGES;
        const InT *xt = sx + (size_t)s * T;
        const int t0 = i * T;
        const int len = EDGE ? min(T, n - t0) : T;
        const BlkTileRec<CarryT> &tr = mi.tr[i & 1];

        // ---- tile mask words ------------------------------------------------------------------
        BlkTileWords tw;
        {
            const int wa = i * NW + lane, wb = wa + 32;
            tw.w0 = (lane < NW && wa < nwords) ? smask[wa] : 0u;
            tw.w1 = (NW > 32 && lane + 32 < NW && wb < nwords) ? smask[wb] : 0u;
            tw.nz = (unsigned long long)__ballot_sync(0xffffffffu, tw.w0 != 0u);
            if (NW > 32) tw.nz |= (unsigned long long)__ballot_sync(0xffffffffu, tw.w1 != 0u) << 32;
        }
        mbar_wait(full0 + 8 * s, (i / STAGES) & 1);

        // ---- one warp per tile: the record of tile i + 1 (its loads are in flight while the span is processed) ----
        const bool rec_warp = (warp == i % WARPS) && (i + 1 < tiles);
        int nq0 = 0, nq1 = 0, nr0 = 0, nr1 = 0;
        CarryT nqx0 = (CarryT)0, nqx1 = (CarryT)0, nrx0 = (CarryT)0, nrx1 = (CarryT)0;
        if (rec_warp) {
            int a, b;
            blk_prev2(tw, 64, a, b);
            if (a >= 0) {
                nq0 = t0 + a;
                nqx0 = (CarryT)xt[a];
                if (b >= 0) {
                    nq1 = t0 + b;
                    nqx1 = (CarryT)xt[b];
                } else {
                    nq1 = tr.qpos[0];
                    nqx1 = tr.qx[0];
                }
            } else {
                nq0 = tr.qpos[0];
                nqx0 = tr.qx[0];
                nq1 = tr.qpos[1];
                nqx1 = tr.qx[1];
            }
            const int lim = min((i + 2) * T, n - 1);           // tile i + 1 ends here
            const int o0 = tr.rpos[0], o1 = tr.rpos[1];
            const CarryT ox0 = tr.rx[0], ox1 = tr.rx[1];
            if (o0 >= lim) {                                   // both still lie beyond the next tile
                nr0 = o0; nrx0 = ox0; nr1 = o1; nrx1 = ox1;
            } else if (o1 >= lim) {
                nr0 = o1; nrx0 = ox1;
                nr1 = (o1 == kNoKnot) ? kNoKnot : blk_next_knot(smask, nwords, o1 + 1, n, lane);
                nrx1 = (nr1 == kNoKnot) ? (CarryT)0 : (CarryT)x[nr1];
            } else {
                nr0 = blk_next_knot(smask, nwords, lim, n, lane);
                nr1 = (nr0 == kNoKnot) ? kNoKnot : blk_next_knot(smask, nwords, nr0 + 1, n, lane);
                nrx0 = (nr0 == kNoKnot) ? (CarryT)0 : (CarryT)x[nr0];
                nrx1 = (nr1 == kNoKnot) ? (CarryT)0 : (CarryT)x[nr1];
            }
        }

        const bool span_live = !EDGE || span_rel < len;
        const int a_glob = t0 + span_rel, b_glob = a_glob + SPAN;
        const bool have_right = span_live && (!EDGE || b_glob <= n - 1);
        CarryT xright = (CarryT)0;
        if (have_right) {
            if (warp < WARPS - 1) {
                xright = (CarryT)xt[span_rel + SPAN];
            } else {
                // first sample of the NEXT tile: its load was issued STAGES-1 tiles ago
                const int s2 = (i + 1) % STAGES;
                mbar_wait(full0 + 8 * s2, ((i + 1) / STAGES) & 1);
                xright = (CarryT)sx[(size_t)s2 * T];
            }
        }

        CarryT xv[IT], bv[IT];
#pragma unroll
        for (int k = 0; k < IT; ++k) xv[k] = bv[k] = (CarryT)0;
        const int nvalid = EDGE ? max(0, min(IT, len - lane_rel)) : IT;
        if (!EDGE || nvalid > 0) blk_load<IT, CarryT>(xt + lane_rel, xv);

        CarryT bleft = (CarryT)0, bright = (CarryT)0;
        if (span_live) {
            const unsigned spanbits = (unsigned)((tw.nz >> sw) & (unsigned long long)ITMASK);
            const bool knot_free = (spanbits == 0u);
            bool cached = knot_free && c_end >= b_glob;
            unsigned bm = 0u;              // this lane's IT flag bits
            int pre = 0, m = 0;            // knots of the span before this lane / in the span
            int r0pos = c_end;
            if (!cached) {
                // ---- knot list of the span: entry 0 = q1, 1 = q0, 2..m+1 = knots inside, m+2 = r0, m+3 = r1 ----
                if (!knot_free) {
                    const unsigned char *mb = reinterpret_cast<const unsigned char *>(smask) + ((a_glob + lane * IT) >> 3);
                    if (!EDGE || a_glob + lane * IT < nmask_bits)
                        bm = (IT == 8) ? (unsigned)mb[0] : (((unsigned)mb[0] >> ((lane & 1) * 4)) & 0xfu);
                    int incl = __popc(bm);
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int t = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += t;
                    }
                    pre = incl - __popc(bm);
                    m = __shfl_sync(0xffffffffu, incl, 31);
                }
                int qa, qb, ra, rb;
                blk_prev2(tw, sw, qa, qb);
                blk_next2(tw, ew, ra, rb);
                if (lane == 0) {
                    int q0p, q1p, r0p, r1p;
                    CarryT q0x, q1x, r0x, r1x;
                    if (qa >= 0) {
                        q0p = t0 + qa; q0x = (CarryT)xt[qa];
                        if (qb >= 0) { q1p = t0 + qb; q1x = (CarryT)xt[qb]; }
                        else { q1p = tr.qpos[0]; q1x = tr.qx[0]; }
                    } else {
                        q0p = tr.qpos[0]; q0x = tr.qx[0];
                        q1p = tr.qpos[1]; q1x = tr.qx[1];
                    }
                    if (ra >= 0) {
                        r0p = t0 + ra; r0x = (CarryT)xt[ra];
                        if (rb >= 0) { r1p = t0 + rb; r1x = (CarryT)xt[rb]; }
                        else { r1p = tr.rpos[0]; r1x = tr.rx[0]; }
                    } else {
                        r0p = tr.rpos[0]; r0x = tr.rx[0];
                        r1p = tr.rpos[1]; r1x = tr.rx[1];
                    }
                    ls[0].pos = q1p; xk[0] = q1x;
                    ls[1].pos = q0p; xk[1] = q0x;
                    ls[m + 2].pos = r0p; xk[m + 2] = r0x;
                    ls[m + 3].pos = r1p; xk[m + 3] = r1x;
                }
                r0pos = (ra >= 0) ? t0 + ra : tr.rpos[0];
                if (!knot_free) {
#pragma unroll
                    for (int k = 0; k < IT; ++k) {
                        if ((bm >> k) & 1u) {
                            const int en = 2 + pre + __popc(bm & ((1u << k) - 1u));
                            ls[en].pos = a_glob + lane * IT + k;
                            xk[en] = xv[k];
                        }
                    }
                }
                __syncwarp();
                // ---- knot baseline L_j, j = 1..m+2 (ITD.py:100-110) ----
                for (int j = 1 + lane; j <= m + 2; j += 32) {
                    const int pj = ls[j].pos;
                    CarryT L;
                    if (pj == 0) {
                        L = endl0;
                    } else if (pj == n - 1) {
                        L = endl1;
                    } else {
                        const int pm = ls[j - 1].pos, pp = ls[j + 1].pos;
                        const CarryT xm = xk[j - 1], xp = xk[j + 1];
                        const CarryT w = A::ratio(pj - pm, pp - pm);
                        const CarryT d = A::sub(xp, xm);
                        const CarryT qq = A::add(xm, A::mul(w, d));
                        L = A::add(A::mul((CarryT)0.5, qq), A::mul((CarryT)0.5, xk[j]));
                    }
                    ls[j].L = L;
                }
                __syncwarp();
                // ---- slopes of segments j = 1..m+1 (ITD.py:116); overwrites the positions ----
                for (int j = 1 + lane; j <= m + 1; j += 32) {
                    const CarryT den = A::sub(xk[j + 1], xk[j]);
                    const CarryT sl = A::div(A::sub(ls[j + 1].L, ls[j].L), den);
                    zero_dx |= (den == (CarryT)0);
                    ls[j].s = sl;
                }
                __syncwarp();
                if (knot_free) {
                    cL = ls[1].L; cs = ls[1].s; cX = xk[1]; cLn = ls[2].L;
                    c_end = r0pos;
                    cached = true;
                } else {
                    c_end = -1;
                }
            }

            // ---- B for the span (ITD.py:112-117) + one halo sample each side ----
            const CarryT xl = (warp > 0) ? (CarryT)xt[span_rel - 1] : (CarryT)0;
            if (cached) {
#pragma unroll
                for (int k = 0; k < IT; ++k) bv[k] = A::add(cL, A::mul(cs, A::sub(xv[k], cX)));
                bleft = (warp == 0) ? mi.carry_b[(i + 1) & 1] : A::add(cL, A::mul(cs, A::sub(xl, cX)));
                if (have_right)
                    bright = (b_glob == c_end) ? cLn : A::add(cL, A::mul(cs, A::sub(xright, cX)));
            } else {
                int cnt = 1 + pre;
#pragma unroll
                for (int k = 0; k < IT; ++k) {
                    cnt += (int)((bm >> k) & 1u);
                    const LS q = ls[cnt];
                    bv[k] = A::add(q.L, A::mul(q.s, A::sub(xv[k], xk[cnt])));
                }
                if (warp == 0) {
                    bleft = mi.carry_b[(i + 1) & 1];
                } else {
                    const LS q = ls[1];
                    bleft = A::add(q.L, A::mul(q.s, A::sub(xl, xk[1])));
                }
                if (have_right) {
                    if (b_glob == r0pos) {
                        bright = ls[m + 2].L;            // x[b] is the knot itself: B = L_k (+ slope * 0)
                    } else {
                        const LS q = ls[m + 1];
                        bright = A::add(q.L, A::mul(q.s, A::sub(xright, xk[m + 1])));
                    }
                }
            }
            if (EDGE) {
                // ITD.py:112: the last sample of every baseline stays 0
                const int kz = n - 1 - (t0 + lane_rel);
#pragma unroll
                for (int k = 0; k < IT; ++k)
                    if (k == kz) bv[k] = (CarryT)0;
                if (b_glob == n - 1) bright = (CarryT)0;
            }
        }

        // ---- outputs: R (row e), B (carry and, on request, the baseline row) ----
        if (!EDGE || nvalid > 0) {
            CarryT rv[IT];
#pragma unroll
            for (int k = 0; k < IT; ++k) {
                const CarryT rr = A::sub(xv[k], bv[k]);                 // ITD.py:119
                rv[k] = LAST ? A::add(rr, bv[k]) : rr;                  // ITD.py:420
            }
            blk_store<IT, CarryT>(rot, rv, nvalid);
            blk_store<IT, CarryT>(carry, bv, nvalid);
            if (BAS) {
                if (LAST) {
#pragma unroll
                    for (int k = 0; k < IT; ++k) rv[k] = (CarryT)0;     // ITD.py:424
                    blk_store<IT, CarryT>(bas, rv, nvalid);
                } else {
                    blk_store<IT, CarryT>(bas, bv, nvalid);
                }
            }
        }
        rot += T;
        carry += T;
        if (BAS) bas += T;

        // ---- extrema of B = the next level's knot flags (ITD.py:44-59 on B and -B) ----
        {
            CarryT give = bv[0];
            CarryT nxt = __shfl_down_sync(0xffffffffu, give, 1);
            if (lane == 31) nxt = bright;
            unsigned lt = 0u, gt = 0u;
#pragma unroll
            for (int k = 0; k < IT; ++k) {
                const CarryT cur = bv[k];
                const CarryT nx = (k + 1 < IT) ? bv[(k + 1 < IT) ? k + 1 : k] : nxt;
                lt |= (cur < nx) ? (1u << k) : 0u;
                gt |= (cur > nx) ? (1u << k) : 0u;
            }
            // the pair (previous sample, my first sample): the previous lane's last pair
            unsigned pairs = ((lt >> (IT - 1)) & 1u) | (((gt >> (IT - 1)) & 1u) << 1);
            unsigned in = __shfl_up_sync(0xffffffffu, pairs, 1);
            if (lane == 0) in = ((bleft < bv[0]) ? 1u : 0u) | ((bleft > bv[0]) ? 2u : 0u);
            // valley: !(b[t-1] < b[t]) && b[t] < b[t+1];  peak: !(b[t-1] > b[t]) && b[t] > b[t+1]
            unsigned f = ((~((lt << 1) | (in & 1u)) & lt) | (~((gt << 1) | (in >> 1)) & gt)) & ITMASK;
            const int tg = t0 + lane_rel;                                // global index of bit 0
            if (EDGE) {
                if (tg == 0) f &= ~1u;                                    // ITD.py:70-73
                const int lastbit = n - 2 - tg;                           // highest valid bit
                f = (lastbit < 0) ? 0u : ((lastbit >= IT - 1) ? f : (f & (ITMASK >> (IT - 1 - lastbit))));
            }
            mycnt += __popc(f);
            if (IT == 8) {
                if (!EDGE || tg < nmask_bits) nmaskb[tg >> 3] = (unsigned char)f;
            } else {
                const unsigned hi = __shfl_down_sync(0xffffffffu, f, 1);
                if (!(lane & 1) && (!EDGE || tg < nmask_bits)) nmaskb[tg >> 3] = (unsigned char)(f | (hi << 4));
            }
        }
        if (warp == WARPS - 1 && lane == 31) mi.carry_b[i & 1] = bv[IT - 1];
        if (rec_warp && lane == 0) {
            BlkTileRec<CarryT> &nt = mi.tr[(i + 1) & 1];
            nt.qpos[0] = nq0; nt.qx[0] = nqx0;
            nt.qpos[1] = nq1; nt.qx[1] = nqx1;
            nt.rpos[0] = nr0; nt.rx[0] = nrx0;
            nt.rpos[1] = nr1; nt.rx[1] = nrx1;
        }
        // every warp is past its reads of stage s: refill it with the tile STAGES ahead
        named_barrier_sync(1, WARPS * 32);
        if (tid == 0 && i + STAGES < tiles) issue_tile(i + STAGES);
    };

    for (int i = 0; i < tiles; ++i) {
        if (i == 0 || i == tiles - 1)
            tile_body(std::true_type{}, i);
        else
            tile_body(std::false_type{}, i);
    }

    // ---- stop test on the device (ITD.py:400-404, :418) --------------------------------------
    if (zero_dx) atomicOr(p.status + sig, kStZeroDx);
    {
        const int wsum = __reduce_add_sync(0xffffffffu, mycnt);
        if (lane == 0) mi.cnt[warp] = wsum;
    }
    __syncthreads();
    if (tid == 0) {
        int Kn = 0;
        for (int w = 0; w < WARPS; ++w) Kn += mi.cnt[w];
        p.next.kcount[sig] = Kn;
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

}  // namespace pyitd
