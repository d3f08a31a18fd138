// itd_stream.cuh -- the batched-channel level kernel: ONE CTA PER SIGNAL, tiles walked in order.
//
// Same arithmetic and the same HBM data structures as level_kernel (itd_kernels.cuh); the
// difference is how a signal is moved through the SM:
//
//   * a producer warp streams each tile's samples, knot-flag words and knot-table slice into a
//     shared-memory ring with TMA bulk copies (cp.async.bulk + mbarrier complete_tx), several
//     tiles ahead of the math;
//   * eight consumer warps each own a contiguous 32*ITEMS-sample span of the tile.  A warp derives
//     its segment ids from the stored flag words (no stencil re-run, no block scan), evaluates the
//     knot baseline / slopes for exactly the knots its span touches (ITD.py:106-110,116) in warp-
//     private scratch, evaluates B and R (ITD.py:115-119), streams them out, and finds the next
//     level's knots from B with shuffles + ballots;
//   * because the tiles of a signal are visited in order by one CTA, the running knot count is a
//     register: no look-back chain, and exactly ONE block barrier per tile (the exchange of the
//     per-warp new-knot counts).
//
// Used when the batch has enough signals to fill the GPU with one CTA each; few long signals use
// the multi-CTA look-back kernel instead.
#pragma once

#include "itd_kernels.cuh"

namespace pyitd {

// ---------------------------------------------------------------------------------------------
// mbarrier / TMA-bulk PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) {
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// global -> shared bulk copy; dst, src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, unsigned bytes,
                                            unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
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
template <typename InT, typename CarryT, int WARPS, int ITEMS, int STAGES>
struct StreamSmem {
    static constexpr int T = WARPS * 32 * ITEMS;
    static constexpr int KC = T + 16;           // knot slice capacity (T + 5, start aligned down to 4)
    static constexpr int SPAN = 32 * ITEMS;
    static constexpr int SC = SPAN + 8;         // per-warp knot scratch
    static constexpr int MAX_TILES = 2048;
    struct Stage {
        alignas(16) InT x[T];
        alignas(16) unsigned mask[(T / 32 + 3) & ~3];
        alignas(16) int tau[KC];
        alignas(16) CarryT xk[KC];
    };
    Stage stage[STAGES];
    alignas(8) unsigned long long full[STAGES];
    alignas(8) unsigned long long empty[STAGES];
    CarryT kL[WARPS][SC];
    CarryT kS[WARPS][SC];
    int tbase[MAX_TILES + 1];
    int cnt[2][WARPS];
    CarryT carry_b[2];                           // B of the previous tile's last sample (by tile parity)
    CarryT endl[2];
};

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <typename InT, typename CarryT, typename OutT, int WARPS, int ITEMS, int STAGES>
__global__ void __launch_bounds__((WARPS + 1) * 32) level_stream_kernel(const LevelParams p) {
    using A = Arith<CarryT>;
    using Smem = StreamSmem<InT, CarryT, WARPS, ITEMS, STAGES>;
    constexpr int T = Smem::T;
    constexpr int SPAN = Smem::SPAN;
    constexpr int NWORDS = T / 32;
    static_assert(NWORDS <= 32, "one flag word per lane");
    extern __shared__ __align__(128) unsigned char smem_stream_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_stream_raw);

    const int sig = blockIdx.x;
    const int n = p.n, e = p.e, tiles = p.tiles;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long row_off = (long long)sig * p.out_sig_stride;

    // ---- signals that already stopped (same rules as level_kernel) ---------------------------
    const int se = p.stop_e[sig];
    if (e > se) {
        OutT *rot = reinterpret_cast<OutT *>(p.rot) + row_off;
        OutT *bas = p.bas ? reinterpret_cast<OutT *>(p.bas) + row_off : nullptr;
        if (e == se + 1 && p.stop_kind[sig] == kStopKnots) {
            const CarryT *src = reinterpret_cast<const CarryT *>(p.fix_src) + (long long)sig * n;
            for (int t = tid; t < n; t += blockDim.x) {
                rot[(long long)se * n + t] = (se == 0) ? (OutT)0 : (OutT)src[t];
                if (bas && (p.opts & kOptZeroTail)) bas[(long long)se * n + t] = (OutT)0;
            }
        }
        if ((p.opts & kOptZeroTail) && e < p.rows) {
            for (int t = tid; t < n; t += blockDim.x) {
                rot[(long long)e * n + t] = (OutT)0;
                if (bas) bas[(long long)e * n + t] = (OutT)0;
            }
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
            for (int s = 0; s < STAGES; ++s) {
                mbar_init(&sm.full[s], 1);
                mbar_init(&sm.empty[s], WARPS);
            }
            mbar_fence_init();
        }
    }
    __syncthreads();

    const InT *x = reinterpret_cast<const InT *>(p.in) + (long long)sig * n;
    const int *gtau = p.cur.tau + (long long)sig * p.cur.kstride;
    const CarryT *gxk = reinterpret_cast<const CarryT *>(p.cur.xk) + (long long)sig * p.cur.kstride;
    const unsigned *gmask_in = p.cur.mask + (long long)sig * p.cur.mstride;

    // =========================================================================================
    // producer warp: TMA bulk loads, STAGES tiles deep
    // =========================================================================================
    if (warp == WARPS) {
        if (lane == 0) {
            for (int i = 0; i < tiles; ++i) {
                const int s = i % STAGES;
                mbar_wait(&sm.empty[s], ((i / STAGES) & 1) ^ 1);
                typename Smem::Stage &st = sm.stage[s];
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
                mbar_arrive_expect_tx(&sm.full[s], bx + bm + bt + bk);
                tma_load_1d(st.x, x + t0, bx, &sm.full[s]);
                tma_load_1d(st.mask, gmask_in + (t0 >> 5), bm, &sm.full[s]);
                tma_load_1d(st.tau, gtau + lo, bt, &sm.full[s]);
                tma_load_1d(st.xk, gxk + lo, bk, &sm.full[s]);
            }
        }
        return;
    }

    // =========================================================================================
    // consumer warps
    // =========================================================================================
    OutT *rot = reinterpret_cast<OutT *>(p.rot) + row_off + (long long)e * n;
    OutT *bas = p.bas ? reinterpret_cast<OutT *>(p.bas) + row_off + (long long)e * n : nullptr;
    CarryT *carry = reinterpret_cast<CarryT *>(p.carry_out) + (long long)sig * n;
    int *ntau = p.next.tau + (long long)sig * p.next.kstride;
    CarryT *nxk = reinterpret_cast<CarryT *>(p.next.xk) + (long long)sig * p.next.kstride;
    unsigned *nmask = p.next.mask + (long long)sig * p.next.mstride;
    int *ntbase = p.next.tbase + (long long)sig * (tiles + 1);
    CarryT *nendl = reinterpret_cast<CarryT *>(p.next.endl) + 2ll * sig;
    const bool last_level = (e == p.emax);
    CarryT *kL = sm.kL[warp];
    CarryT *kS = sm.kS[warp];
    const unsigned lt_mask = (1u << lane) - 1u;
    const unsigned le_mask = 0xffffffffu >> (31 - lane);

    int run_total = 0;          // new-level knots found in earlier tiles (identical in every warp)
    bool zero_dx = false;

    for (int i = 0; i < tiles; ++i) {
        const int s = i % STAGES;
        typename Smem::Stage &st = sm.stage[s];
        const int t0 = i * T;
        const int len = min(T, n - t0);
        const int span0 = warp * SPAN;                    // first sample of this warp's span (in tile)
        const int kb = sm.tbase[i], cnt = sm.tbase[i + 1] - kb;
        const int lo = max(kb - 1, 0) & ~3;
        mbar_wait(&sm.full[s], (i / STAGES) & 1);

        // ---- A. segment bases from the stored flag words --------------------------------------
        const int nwords = (len + 31) >> 5;
        const unsigned word = (lane < nwords) ? st.mask[lane] : 0u;
        const int pc = __popc(word);
        int incl = pc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const int excl = incl - pc;
        unsigned mw[ITEMS];
        int wpre[ITEMS];
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) {
            mw[r] = shfl_idx(word, warp * ITEMS + r);
            wpre[r] = shfl_idx(excl, warp * ITEMS + r);
        }
        const int wb = kb + wpre[0];                                  // knots before the span = seg(span0 - 1)
        const int wcnt = wpre[ITEMS - 1] + __popc(mw[ITEMS - 1]) - wpre[0];   // knots inside the span
        const bool span_live = span0 < len;

        // right-halo sample (first sample after the span): value, flag, availability
        const int tend = t0 + span0 + SPAN;                           // global index of that sample
        bool have_right = span_live && (tend <= n - 1);
        CarryT xright = (CarryT)0;
        int fright = 0;
        if (have_right) {
            if (warp < WARPS - 1) {
                xright = (CarryT)st.x[span0 + SPAN];
                fright = (int)(shfl_idx(word, (warp + 1) * ITEMS) & 1u);
            } else {
                // first sample of the NEXT tile: the producer is ahead, wait for its stage
                const int s2 = (i + 1) % STAGES;
                mbar_wait(&sm.full[s2], ((i + 1) / STAGES) & 1);
                xright = (CarryT)sm.stage[s2].x[0];
                fright = (int)(sm.stage[s2].mask[0] & 1u);
            }
        }

        // ---- B. knot baseline + slopes for the knots this span touches (warp-private) ----------
        // L for k in [wb, wb + wcnt + 2], slope for segments [wb, wb + wcnt + 1], clipped to the table
        if (span_live) {
            const int nl = min(wcnt + 3, K + 2 - wb);
            for (int j = lane; j < nl; j += 32) {
                const int k = wb + j;
                const int q = k - lo;
                CarryT L;
                if (k == 0) {
                    L = sm.endl[0];
                } else if (k == K + 1) {
                    L = sm.endl[1];
                } else {
                    const CarryT w = A::ratio(st.tau[q] - st.tau[q - 1], st.tau[q + 1] - st.tau[q - 1]);
                    const CarryT d = A::sub(st.xk[q + 1], st.xk[q - 1]);
                    const CarryT qq = A::add(st.xk[q - 1], A::mul(w, d));
                    L = A::add(A::mul((CarryT)0.5, qq), A::mul((CarryT)0.5, st.xk[q]));
                }
                kL[j] = L;
            }
            __syncwarp();
            const int ns = min(wcnt + 2, K + 1 - wb);
            for (int j = lane; j < ns; j += 32) {
                const int q = wb + j - lo;
                const CarryT den = A::sub(st.xk[q + 1], st.xk[q]);
                kS[j] = A::div(A::sub(kL[j + 1], kL[j]), den);
                zero_dx |= (den == (CarryT)0);
            }
            __syncwarp();
        }

        // ---- C. B, R for the span (+ one halo sample each side) --------------------------------
        CarryT b[ITEMS];
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) {
            const int jt = span0 + r * 32 + lane;
            const int t = t0 + jt;
            const CarryT xv = (jt < len) ? (CarryT)st.x[jt] : (CarryT)0;
            const int j = min(wpre[r] - wpre[0] + __popc(mw[r] & le_mask), K - wb);
            CarryT bv = (CarryT)0;
            if (jt < len) {
                bv = A::add(kL[j], A::mul(kS[j], A::sub(xv, st.xk[wb + j - lo])));
                if (t == n - 1) bv = (CarryT)0;                       // ITD.py:112
                const CarryT rr = A::sub(xv, bv);
                rot[t] = (OutT)(last_level ? A::add(rr, bv) : rr);    // ITD.py:119 / :420
                carry[t] = bv;
                if (bas) bas[t] = last_level ? (OutT)0 : (OutT)bv;    // ITD.py:424
                if (t == n - 2) nendl[1] = mean2<CarryT>(bv, (CarryT)0);
            }
            b[r] = bv;
        }
        // left halo B[span0 - 1]: previous warp's last sample (same tile) or the previous tile's
        CarryT bleft = (CarryT)0;
        if (span_live) {
            if (warp == 0) {
                bleft = sm.carry_b[(i + 1) & 1];
            } else {
                const CarryT xl = (CarryT)st.x[span0 - 1];
                const int j = min(0, K - wb);
                bleft = A::add(kL[j], A::mul(kS[j], A::sub(xl, st.xk[wb + j - lo])));
            }
        }
        CarryT bright = (CarryT)0;
        if (have_right && tend < n - 1) {
            const int j = min(wcnt + fright, K - wb);
            bright = A::add(kL[j], A::mul(kS[j], A::sub(xright, st.xk[wb + j - lo])));
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[s]);                     // stage consumed by this warp

        // ---- D. extrema of B: next level's flag words -----------------------------------------
        unsigned fw[ITEMS];
        int newc = 0;
        {
            unsigned lt_in, gt_in;
            {
                const CarryT b0 = shfl_idx(b[0], 0);
                lt_in = (bleft < b0) ? 1u : 0u;
                gt_in = (bleft > b0) ? 1u : 0u;
            }
#pragma unroll
            for (int r = 0; r < ITEMS; ++r) {
                CarryT nx = __shfl_down_sync(0xffffffffu, b[r], 1);
                const CarryT wrap = (r + 1 < ITEMS) ? shfl_idx(b[(r + 1 < ITEMS) ? r + 1 : r], 0) : bright;
                if (lane == 31) nx = wrap;
                const unsigned LT = __ballot_sync(0xffffffffu, b[r] < nx);
                const unsigned GT = __ballot_sync(0xffffffffu, b[r] > nx);
                unsigned f = (~((LT << 1) | lt_in) & LT) | (~((GT << 1) | gt_in) & GT);
                lt_in = LT >> 31;
                gt_in = GT >> 31;
                // valid positions: 1 <= t <= n-2
                const int tw = t0 + span0 + r * 32;                   // global index of bit 0
                if (tw == 0) f &= ~1u;
                const int lastbit = n - 2 - tw;                       // highest valid bit
                f = (lastbit < 0) ? 0u : ((lastbit >= 31) ? f : (f & (0xffffffffu >> (31 - lastbit))));
                fw[r] = f;
                newc += __popc(f);
            }
        }
        if (lane == 0) sm.cnt[i & 1][warp] = newc;
        if (lane < ITEMS && span0 + lane * 32 < len) {
            unsigned v = fw[0];
#pragma unroll
            for (int r = 1; r < ITEMS; ++r) v = (lane == r) ? fw[r] : v;
            nmask[(t0 + span0) / 32 + lane] = v;
        }
        if (warp == WARPS - 1 && lane == 31) sm.carry_b[i & 1] = b[ITEMS - 1];
        if (i == 0 && warp == 0) {
            const CarryT b1 = shfl_idx(b[0], 1);
            if (lane == 0) {
                ntau[0] = 0;
                nxk[0] = b[0];
                nendl[0] = mean2<CarryT>(b[0], b1);
            }
        }

        // ---- E. the one block barrier per tile -----------------------------------------------
        named_barrier_sync(1, WARPS * 32);

        // ---- F/G. compact the new knots -------------------------------------------------------
        int pre = run_total, tot = 0;
#pragma unroll
        for (int w2 = 0; w2 < WARPS; ++w2) {
            const int c = sm.cnt[i & 1][w2];
            pre += (w2 < warp) ? c : 0;
            tot += c;
        }
        if (warp == 0 && lane == 0) ntbase[i] = run_total;
        run_total += tot;
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) {
            if ((fw[r] >> lane) & 1u) {
                const int rank = pre + __popc(fw[r] & lt_mask);
                ntau[1 + rank] = t0 + span0 + r * 32 + lane;
                nxk[1 + rank] = b[r];
            }
            pre += __popc(fw[r]);
        }
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
        } else if (last_level) {                                      // ITD.py:418
            p.stop_kind[sig] = kStopIter;
            p.n_rows[sig] = e + 1;
            p.stop_e[sig] = e;
        }
    }
}

}  // namespace pyitd
