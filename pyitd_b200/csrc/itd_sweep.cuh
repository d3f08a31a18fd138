// itd_sweep.cuh -- the batched-channel path of round 2: the WHOLE decomposition of a batch in ONE launch.
//
// A persistent grid (as many CTAs as fit) pulls work items (stage, signal) from a ticket counter, stage-major:
// stage -1 is the extrema-compaction pass over the raw input (ITD.py:87-98), stage e >= 0 is extraction e
// (ITD.py:79-121 applied to X_e, plus the driver's stop test ITD.py:400-404, which IS the next level's knot detection).
// The item (e, s) waits for (e-1, s) through a per-signal counter (release / acquire); because tickets are handed out
// in order, the item it waits for was taken earlier by a CTA that is running, so the wait cannot deadlock, and with
// more signals than resident CTAs it never actually spins.  There are no launch boundaries between levels: no partial
// last wave per level, no tail of nearly empty launches once most signals have stopped, no host involvement.
//
// Inside an item the eight warps of the CTA do not talk to each other while they stream:
//
//   * warp w owns the w-th contiguous REGION of the signal (1/8 of its 128-sample spans) and walks it span by span.
//     Samples come straight from global memory into registers (coalesced 8-byte loads, the next span's loads issued
//     before the current span's arithmetic, an L2 prefetch a few spans further ahead); R and B go straight back with
//     coalesced stores.  No shared-memory staging of samples, no mbarrier, no block barrier per tile.
//   * the knots a warp finds in B (= the next level's knots) are compacted into the warp's OWN region list
//     (tau, X) in the order found, with the running count in a register: the global rank of a knot is not needed
//     while streaming.  The knot flags are also kept as a bit mask (segment id of a sample = knots at or before it).
//   * the per-knot work of ITD.py:100-110 and :116 -- L_k and the slope of the segment that starts at knot k --
//     is done ONE THREAD PER KNOT:
//       - few knots (K + 2 <= kSweepCap): by the whole block, once per item, into a shared-memory table
//         {X_k, L_k, s_k} indexed by global rank (three block barriers per signal-level instead of one per tile);
//       - many knots (the first two or three levels): per warp, for as many consecutive spans of its region as fit a
//         warp-private table of 250 records (3 spans on level 0 of the benchmark, 10 on level 1, 30 on level 2), 32
//         knots at a time with every lane busy, from the region's own list, whose two slots before and three slots
//         after the region's knots are filled with the neighbouring regions' knots ("halo") by 40 threads at the start
//         of the item.  The spans of such a chunk then stream exactly like spans of a few-knot level.
//
//   * FUSED PAIRS (sweep_region_fused, predict_pair): two consecutive few-knot extractions e, e + 1 run as ONE pass that
//     reads X_e and writes R_e, R_{e+1}, B_{e+1}; B_e lives in registers only (32 bytes per sample instead of 48).  The
//     pass needs the knots of B_e up front; they are PREDICTED from the knot table without touching the signal -- between
//     two knots of X_e the baseline is a monotone function of a monotone stretch, so its extrema sit at knots of X_e (or
//     at sample n-2, whose right neighbour is forced to 0) unless two neighbouring values are equal -- and the pass CHECKS
//     the prediction on every sample: if a flag word differs the item redoes extraction e on its own.  The predicted knot
//     count (a lower bound of the true one) also decides whether an extraction that is not fused is probed first.
//
// Same arithmetic, same operation order, -fmad=false: bit-identical to the reference in fp64.
#pragma once

#include <type_traits>

#include "itd_kernels.cuh"
#include "itd_stream.cuh"

namespace pyitd {

constexpr int kSweepWarps = 8;
constexpr int kSweepItems = 4;
constexpr int kSweepSpan = 32 * kSweepItems;          // samples per warp iteration
// The knot lists hold (tau_k, X_k).  -DPYITD_SWEEP_TAU_ONLY keeps tau only and gathers X_k = X_e[tau_k] from the input when a
// table is built (-16 B of traffic per knot: the scan stage 0.73 -> 0.57 ms); measured and NOT adopted: the dependent
// loads (tau, then the sample) make every chunk build of a many-knot level and every table build wait twice
// (e = 0 .. 3: 2.10 / 1.47 / 1.26 / 1.52 -> 2.21 / 1.64 / 1.56 / 1.65 ms; the step 12.75 -> 13.4 ms), profiles/r2/README.md.
#ifndef PYITD_SWEEP_TAU_ONLY
#define PYITD_SWEEP_XK_LISTS 1
#endif
#ifndef PYITD_SWEEP_CAP
#define PYITD_SWEEP_CAP 2240
#endif
// shared-memory knot table: K + 2 <= kSweepCap.  Up to 2000 entries the kernel's shared memory is static (48 KB); above, it is
// dynamic (opt-in size): 2240 entries = 54 KB is the most that still lets four CTAs share an SM.
constexpr int kSweepCap = PYITD_SWEEP_CAP;
constexpr bool kSweepDynSmem = (kSweepCap > 2000);
constexpr int kSweepPre = 2, kSweepPost = 3;          // halo slots of a region list
constexpr int kSweepScratch = kSweepCap / kSweepWarps;      // warp-private table entries of a many-knot item (>= span knots + 5)
// fused pairs: for extraction e with kSweepFuseMinA <= K and K + 2 <= kSweepFuseMaxA the knots of its baseline are PREDICTED
// from the table (see sweep_kernel); with at least kSweepFuseMinB of them (and both tables fitting the block's arrays)
// extractions e and e + 1 run as one pass that also checks the prediction
constexpr int kSweepFuseMinA = 12, kSweepFuseMaxA = 1640, kSweepFuseMinB = 4;
constexpr int kSweepFuseMaxLimit = 1780;              // the prediction step handles at most this many knots
constexpr int kSweepProbeKnots = 3;                   // extractions with at most this many knots are probed first
static_assert(kSweepScratch >= kSweepSpan + 5, "a span's knots + 5 must fit a warp table");
enum { kPtrIn = 0, kPtrRot, kPtrBas, kPtrCarry, kPtrGmask, kPtrNmask, kPtrCtau, kPtrCxk, kPtrNtau, kPtrNxk,
       kPtrRot2, kPtrBas2, kPtrGmask2, kSweepPtrs };
constexpr int kSweepDoneAll = 0x0fffffff;             // done[] value (low 28 bits) of a signal that has stopped

struct SweepTable {
    int *tau;            // [S, 8 * rs]  region r at r * rs: kSweepPre halo slots, the region's knots in order, kSweepPost halo slots
    void *xk;            // same layout (carry type): X at the knot
    unsigned *mask;      // [S, mstride] knot flags of the level's input, bit (t & 31) of word (t >> 5)
    int *rcount;         // [S, 8]       knots per region
};

struct SweepParams {
    const void *x;       // [S, N] input type
    // Ping-pong, selected PER SIGNAL by a bit that travels with done[] (0 for extraction 0, flipped by every item that
    // completes an extraction or a fused pair of extractions): the item reads the knots of its input from tab[sel] and the input itself (e >= 1) from
    // carry[sel ^ 1]; it writes the last baseline it computes to carry[sel] and that baseline's knots to tab[sel ^ 1].
    void *carry[2];      // [S, N] carry type
    SweepTable tab[2];
    // fused pair (e, e + 1): the flag words of B_e's knots, in the CTA's own scratch
    unsigned *mid_mask;  // [mid_ctas, mstride]
    int mid_ctas;        // CTAs the scratch was sized for (0: no fused pairs)
    int *stats;          // [3] pairs of extractions fused; pairs not tried after the prediction (too few / too many knots);
                         //     fused passes whose check failed (the extraction was redone on its own)
    void *rot, *bas;     // [S, rows, N] output type; bas may be null
    long long out_sig_stride;
    long long kstride, mstride;
    int *ticket;         // [1]  zeroed before the launch
    int *done;           // [S]  stages completed: 1 after the scan, e + 2 after extraction e; bit 30: the table selector;
                         //      bit 28: a fused pair of this signal failed its check once -- no further pairs are tried
    int *stop_e, *stop_kind, *n_rows, *knot_counts, *status, *input_knots;
    unsigned long long *stage_ns;   // optional [rows + 1]: CTA-nanoseconds spent per stage (index stage + 1)
    int S, n;
    int spans, spw, rs;  // spans of a signal, spans per warp region, slots per region list
    int stage_first, stage_last;   // stages of this launch: -1 (scan) .. emax
    int emax, rows, min_extrema;
    unsigned opts;
    int pf_scan;                   // ... of the input scan inside extraction 0
    int pf_sparse, pf_dense;       // L2 prefetch distance of the sample stream in spans (0: none), few / many knots
    int pf_fused;                  // ... of the pass of a fused pair
    // item order.  0: stage-major tickets (every signal's stage e before any signal's stage e + 1; an item waits for its
    // signal's previous stage through done[]).  1: a ticket is a SIGNAL and the CTA runs all of its stages back to back:
    // the carry, flag words and knot lists it reads were written by the same CTA a few microseconds earlier and are
    // still in L2 -- for short signals (framed audio), where all resident CTAs' working sets fit there.
    int depth_first;
    int fuse;                      // 1: pairs of few-knot extractions run as one item that never stores the baseline between them
    int predict;                   // 1: the knot count of a few-knot extraction's baseline is predicted from its table
    int fuse_min_a, fuse_max_a, fuse_min_b;   // thresholds (kSweepFuse*; experiment hooks)
    int fused_scan;                // 1: no scan stage (stage_first >= 0): extraction 0 finds the input's knots itself (kFirst)
};

template <typename CarryT>
struct SweepSmem {
    // few knots: {X, L, S}[global rank]; many knots: warp w's scratch at w * kSweepScratch in each array
    static constexpr int kCap = kSweepCap;
    CarryT X[kCap], L[kCap], S[kCap];
    int prefix[kSweepWarps + 1];               // knots before each region (prefix[8] = K)
    int prefix2[kSweepWarps + 1];              // fused pair: the same for the knots of B_e
    CarryT endl2[2], endx2[2];                 // ... and its end values
    CarryT xe[4];                              // X_e at samples 0, 1, n-2, n-1
    int pred_bad;                              // fused pair: an extremum of B_e away from the predicted knots was seen
    int cnt[kSweepWarps];                      // next level's knots per region
    CarryT endl[2], endx[2];                   // L_0, L_{K+1} (ITD.py:101-102); X_0 = in[0], X_{K+1} = in[n-1]
    // the item's base pointers (per signal / per row), computed once per item by one thread: the span loop adds a
    // sample or word index to them instead of re-deriving sig * stride + e * n every span
    void *ptr[kSweepPtrs];
    int keep[kSweepWarps][2];                  // extraction 0 without a scan stage: tau of a chunk's last two elements
    int knots_in[kSweepWarps];                 // ... and the input knots each warp found in its own spans
    int ticket, zero_dx, dn;
    // the item's scalars, parked here while the warps stream (reloaded afterwards: no register is held across the span loops)
    int it_e, it_sig, it_sel, it_de, it_dsig, it_flags, it_k2, it_nofuse;
    unsigned long long t_start;
};

__device__ __forceinline__ int ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// data written earlier in this launch by another CTA: read through L2 (the L1 of this SM may hold a stale line)
template <typename T>
__device__ __forceinline__ T ld_cg(const T *p) {
    return __ldcg(p);
}

// ---------------------------------------------------------------------------------------------
// one region of one signal, one stage.  XT = element type of the stage's input (the caller's input type for the scan
// and extraction 0, the carry type afterwards).
//
// The span body exists in three flavours per stage kind, chosen per span with warp-uniform branches:
//   <EDGE, 2>   the spans that hold sample 0 or come within one span of sample n-1: every bound is checked, knot mode
//               taken from the runtime flag;
//   <false, 0>  few knots: records from the block's table;      <false, 1>  many knots: records built per span.
// Non-EDGE spans are followed by a complete span, so their loads, flag words and neighbours need no checks.
// ---------------------------------------------------------------------------------------------
// KIND: kLevel = one extraction; kScan = the extrema-compaction pass over the raw input; kProbe = extraction e of a signal
// with at most kSweepProbeKnots knots, run WITHOUT storing B or the next level's knots: it writes the candidate trend
// row (X_e) into row e and only counts the extrema of B_e.  Two thirds of such extractions are the discarded last one
// (ITD.py:404-411), which then costs one read and one write per sample instead of a full level plus a row copy.
// kFirst = extraction 0 WITHOUT a separate scan stage: each warp finds the knots of the raw input itself (3-point stencil,
// ITD.py:44-59 on x and -x) while it builds the record table of a chunk, straight into shared memory -- the input's
// knot lists (0.56 knots per sample on the benchmark: 12 B written and 12 B read per knot) never exist, and the
// stage that read the input just to find them is gone.
// kRecomp = B_e recomputed into row e + 1 (the trend row when a fused pair's second extraction turns out to be the discarded
// one).
enum { kLevel = 0, kScan = 1, kProbe = 2, kFirst = 3, kRecomp = 5 };
template <typename XT, typename CarryT, typename OutT, int KIND, bool BAS>
__device__ __forceinline__ void sweep_region(const SweepParams &p, SweepSmem<CarryT> &sm, const bool dense,
                                             const bool last, const int K, const int warp, const int lane,
                                             int &region_knots, int &input_knots, bool &zero_dx, bool &bad) {
    using A = Arith<CarryT>;
    constexpr int ITEMS = kSweepItems, SPAN = kSweepSpan;
    constexpr bool SCAN = (KIND == kScan), FIRST = (KIND == kFirst), RECOMP = (KIND == kRecomp);
    constexpr bool PROBE = (KIND == kProbe) || RECOMP;               // no flag words, lists or carry are written
    const int n = p.n;
    const int sp0 = warp * p.spw;
    const int sp1 = min(sp0 + p.spw, p.spans);
    region_knots = 0;
    input_knots = 0;
    if (sp0 >= sp1) return;

    // item base pointers come from shared memory where they are used (no registers held across the span loop)
    auto in_p = [&]() { return reinterpret_cast<const XT *>(sm.ptr[kPtrIn]); };
    auto rot_p = [&]() { return reinterpret_cast<OutT *>(sm.ptr[kPtrRot]); };
    auto bas_p = [&]() { return reinterpret_cast<OutT *>(sm.ptr[kPtrBas]); };
    auto carry_p = [&]() { return reinterpret_cast<CarryT *>(sm.ptr[kPtrCarry]); };
    auto gmask_p = [&]() { return reinterpret_cast<const unsigned *>(sm.ptr[kPtrGmask]); };
    auto nmask_p = [&]() { return reinterpret_cast<unsigned *>(sm.ptr[kPtrNmask]); };
    const int roff = warp * p.rs;                             // this region's list inside the signal's knot arrays

    const int wsc = warp * kSweepScratch;                     // this warp's scratch inside the table arrays
    const int gbase0 = SCAN ? 0 : sm.prefix[warp];            // global rank of the last knot before the region
    const unsigned le_mask = 0xffffffffu >> (31 - lane);

    int pos = 0;                 // knots of this region's list consumed so far
    int npos = 0;                // next-level knots of this region found so far
    CarryT bleft = (CarryT)0;    // value left of the span (previous span's last B, or the previous region's)

    // ---- first span's samples and flag words ---------------------------------------------------
    XT xc[ITEMS];
    uint4 mc = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
        const int t = sp0 * SPAN + r * 32 + lane;
        xc[r] = (t < n) ? ld_cg(in_p() + t) : (XT)0;
    }
    if (!SCAN && !FIRST) mc = ld_cg(reinterpret_cast<const uint4 *>(gmask_p() + sp0 * ITEMS));

    // ---- many knots: warp-private records for a CHUNK of consecutive spans -----------------------------------------
    // As many spans of this region as fit the warp's table (kSweepScratch entries) get their records in one go: the
    // list entries arrive with independent coalesced loads, L and the slopes are computed 32 knots at a time with
    // every lane busy, and the spans of the chunk then stream like spans of a few-knot level.  table[i] = region list
    // slot cbase + i = the knot with global rank g0 + i, g0 = gbase0 + cbase - 1; table[1] is the last knot before the
    // chunk.  L for i in [1, tot+3], slope for i in [1, tot+2]  (ITD.py:100-110, :116)
    int chunk_left = 0, cbase = 0;
    auto build_chunk = [&](const int sp) {
        // plan: the next m <= 32 spans, tot knots, tot + 5 <= kSweepScratch (one span always fits: 128 + 5)
        const int nsp = min(32, sp1 - sp);
        uint4 w4 = make_uint4(0, 0, 0, 0);
        if (lane < nsp) w4 = ld_cg(reinterpret_cast<const uint4 *>(gmask_p()) + sp + lane);
        int pre = __popc(w4.x) + __popc(w4.y) + __popc(w4.z) + __popc(w4.w);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, pre, o);
            if (lane >= o) pre += v;
        }
        const int m = __popc(__ballot_sync(0xffffffffu, lane < nsp && pre + 5 <= kSweepScratch));
        const int tot = __shfl_sync(0xffffffffu, pre, m - 1);
        const int *ctau = reinterpret_cast<const int *>(sm.ptr[kPtrCtau]) + roff + pos;
#ifdef PYITD_SWEEP_XK_LISTS
        const CarryT *cxk = reinterpret_cast<const CarryT *>(sm.ptr[kPtrCxk]) + roff + pos;
#endif
        int *tw = reinterpret_cast<int *>(sm.S + wsc);               // tau words live in S's storage until S is computed
        const int g0 = gbase0 + pos - 1;
        __syncwarp();                                                // the previous chunk's lookups are done
#ifdef PYITD_SWEEP_XK_LISTS
        for (int i = lane; i < tot + 5; i += 32) {
            tw[i] = ld_cg(ctau + i);
            sm.X[wsc + i] = ld_cg(cxk + i);
        }
        if (lane < 12 && pos + tot + 5 + 3 * kSweepScratch < p.rs) {   // the next chunk's entries towards L2
            if (lane < 4) prefetch_l2(ctau + tot + 5 + kSweepScratch / 2 + lane * 32);
            else prefetch_l2(cxk + tot + 5 + kSweepScratch / 2 + (lane - 4) * 16);
        }
#else
        // the lists hold tau only: X_k = X_e[tau_k] is gathered from the input (32 consecutive knots of a many-knot level sit
        // in a handful of 128-byte lines, which the span loop is about to read anyway)
        for (int i = lane; i < tot + 5; i += 32) {
            const int tv = ld_cg(ctau + i);
            tw[i] = tv;
            sm.X[wsc + i] = (CarryT)ld_cg(in_p() + tv);
        }
        if (lane < 4 && pos + tot + 5 + 3 * kSweepScratch < p.rs)      // the next chunk's entries towards L2
            prefetch_l2(ctau + tot + 5 + kSweepScratch / 2 + lane * 32);
#endif
        __syncwarp();
        // the end knots 0 and K+1 (and the unused ranks beyond them) are rare: one warp-uniform test per chunk
        const bool clip = (g0 + 1 <= 0) || (g0 + tot + 3 >= K + 1);
        for (int i = 1 + lane; i <= tot + 3; i += 32) {
            const int g = g0 + i;
            CarryT Lv;
            if (clip && g <= 0) {
                Lv = sm.endl[0];
            } else if (clip && g >= K + 1) {
                Lv = sm.endl[1];
            } else {
                const int tl = tw[i - 1];
                const CarryT w = A::ratio(tw[i] - tl, tw[i + 1] - tl);
                const CarryT xl = sm.X[wsc + i - 1];
                const CarryT d = A::sub(sm.X[wsc + i + 1], xl);
                const CarryT qq = A::add(xl, A::mul(w, d));
                Lv = A::add(A::mul((CarryT)0.5, qq), A::mul((CarryT)0.5, sm.X[wsc + i]));
            }
            sm.L[wsc + i] = Lv;
        }
        __syncwarp();
        for (int i = 1 + lane; i <= tot + 2; i += 32) {
            const int g = g0 + i;
            CarryT sl = (CarryT)0;
            if (!clip || (g >= 0 && g <= K)) {
                const CarryT den = A::sub(sm.X[wsc + i + 1], sm.X[wsc + i]);
                sl = A::div(A::sub(sm.L[wsc + i + 1], sm.L[wsc + i]), den);
                zero_dx |= (den == (CarryT)0);
            }
            sm.S[wsc + i] = sl;
        }
        __syncwarp();
        chunk_left = m;
        cbase = pos;
    };

    // ---- extraction 0 without a scan stage: the chunk's records from the raw input ---------------------------------
    // The element sequence of a signal is: the end knot at sample 0, its interior knots, the end knot at sample n-1.
    // table[0..1] = the two elements before the chunk's first knot, table[2 .. tot+1] = the knots of the chunk's spans,
    // table[tot+2 .. tot+4] = the three elements after them; -1 marks "no such element".  A chunk takes spans while
    // their knots fit the table; the first span that does not (or the spans behind the region) supplies the three
    // elements after it and is scanned again by the next chunk.  Flag words of every scanned span go to the mask array
    // the span loop reads (through L2).
    int ctot = 0;                                                    // knots of the current chunk
    auto scan_span = [&](const int s, CarryT (&v)[ITEMS], unsigned (&fw)[ITEMS], const bool check) -> int {
        const int t0 = s * SPAN;
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) {
            const int t = t0 + r * 32 + lane;
            v[r] = (t < n) ? (CarryT)ld_cg(in_p() + t) : (CarryT)0;
            if (check && t < n) bad |= !isfinite(v[r]);
        }
        const CarryT vl = (t0 > 0) ? (CarryT)ld_cg(in_p() + t0 - 1) : (CarryT)0;
        const CarryT vr = (t0 + SPAN <= n - 1) ? (CarryT)ld_cg(in_p() + t0 + SPAN) : (CarryT)0;
        if (check && lane < SPAN * (int)sizeof(XT) / 128) {          // the scan runs ahead of the span loop: keep DRAM busy
            const long long tp = (long long)t0 + p.pf_scan * SPAN + lane * (128 / (int)sizeof(XT));
            if (tp < n) prefetch_l2(in_p() + tp);
        }
        const int c = span_extrema<true, ITEMS, CarryT>(v, vl, vr, lane, t0, n, fw);
        if (lane < ITEMS) {
            unsigned w = fw[0];
#pragma unroll
            for (int r = 1; r < ITEMS; ++r) w = (lane == r) ? fw[r] : w;
            __stwb(reinterpret_cast<unsigned *>(sm.ptr[kPtrGmask]) + s * ITEMS + lane, w);
        }
        return c;
    };
    auto build_chunk_first = [&](const int sp) {
        int *tw = reinterpret_cast<int *>(sm.S + wsc);               // tau words live in S's storage until S is computed
        CarryT *tx = sm.X + wsc;
        const unsigned lt_mask = le_mask >> 1;
        __syncwarp();                                                // the previous chunk's lookups are done
        // ---- the two elements before the chunk
        if (sp == sp0) {
            int found = 0;
            for (int s = sp0 - 1; s >= 0 && found < 2; --s) {
                CarryT v[ITEMS];
                unsigned fw[ITEMS];
                const int c = scan_span(s, v, fw, false);
                if (c) {
                    int above = 0;                                   // knots of the span in higher words
#pragma unroll
                    for (int r = ITEMS - 1; r >= 0; --r) {
                        if ((fw[r] >> lane) & 1u) {
                            const int rtop = above + __popc(fw[r] & ~le_mask);      // knots of the span behind mine
                            if (rtop < 2 - found) {
                                tw[1 - found - rtop] = s * SPAN + r * 32 + lane;
                                tx[1 - found - rtop] = v[r];
                            }
                        }
                        above += __popc(fw[r]);
                    }
                    found += min(c, 2 - found);
                }
            }
            if (found < 2 && lane == 0) {                            // ran into the start of the signal
                tw[1 - found] = 0;
                tx[1 - found] = sm.endx[0];
                if (found == 0) {
                    tw[0] = -1;
                    tx[0] = (CarryT)0;
                }
            }
        } else {
            const int t0v = sm.keep[warp][0], t1v = sm.keep[warp][1];   // (the slopes have overwritten the tau words)
            const CarryT x0v = tx[ctot], x1v = tx[ctot + 1];
            __syncwarp();
            if (lane == 0) {
                tw[0] = t0v;
                tw[1] = t1v;
                tx[0] = x0v;
                tx[1] = x1v;
            }
        }
        // ---- the chunk's spans, then the three elements after them
        int tot = 0, m = 0, post = 0;
        for (int s = sp; s < p.spans; ++s) {
            CarryT v[ITEMS];
            unsigned fw[ITEMS];
            const int c = scan_span(s, v, fw, s < sp1);
            const bool own = (post == 0) && (s < sp1) && (tot + c + 5 <= kSweepScratch);
            const int room = own ? c : 3 - post;                     // knots of this span that enter the table
            if (c) {
                int pre = 0;
#pragma unroll
                for (int r = 0; r < ITEMS; ++r) {
                    if ((fw[r] >> lane) & 1u) {
                        const int rank = pre + __popc(fw[r] & lt_mask);
                        if (rank < room) {
                            tw[2 + tot + (own ? 0 : post) + rank] = s * SPAN + r * 32 + lane;
                            tx[2 + tot + (own ? 0 : post) + rank] = v[r];
                        }
                    }
                    pre += __popc(fw[r]);
                }
            }
            if (own) {
                tot += c;
                ++m;
            } else {
                post += min(c, 3 - post);
                if (post >= 3) break;
            }
        }
        if (post < 3 && lane == 0) {                                 // ran into the end of the signal
            tw[2 + tot + post] = n - 1;
            tx[2 + tot + post] = sm.endx[1];
            for (int q = post + 1; q < 3; ++q) {
                tw[2 + tot + q] = -1;
                tx[2 + tot + q] = (CarryT)0;
            }
        }
        __syncwarp();
        // ---- L for entries [1, tot+3], slopes for [1, tot+2]  (ITD.py:100-110, :116).  The slopes overwrite the tau
        // words, so which entries start a real segment is noted (one bit per round) while tau is still intact.
        unsigned seg_ok = 0;
        for (int i = 1 + lane, k = 0; i <= tot + 3; i += 32, ++k) {
            const int ti = tw[i];
            CarryT Lv = (CarryT)0;
            if (ti == 0) {
                Lv = sm.endl[0];
            } else if (ti == n - 1) {
                Lv = sm.endl[1];
            } else if (ti > 0) {
                const int tl = tw[i - 1];
                const CarryT w = A::ratio(ti - tl, tw[i + 1] - tl);
                const CarryT xl = tx[i - 1];
                const CarryT d = A::sub(tx[i + 1], xl);
                const CarryT qq = A::add(xl, A::mul(w, d));
                Lv = A::add(A::mul((CarryT)0.5, qq), A::mul((CarryT)0.5, tx[i]));
            }
            if (ti >= 0 && ti != n - 1 && i <= tot + 2 && tw[i + 1] >= 0) seg_ok |= 1u << k;
            sm.L[wsc + i] = Lv;
        }
        if (lane < 2) sm.keep[warp][lane] = tw[tot + lane];          // the next chunk's first two entries
        __syncwarp();
        for (int i = 1 + lane, k = 0; i <= tot + 2; i += 32, ++k) {
            CarryT sl = (CarryT)0;
            if ((seg_ok >> k) & 1u) {
                const CarryT den = A::sub(tx[i + 1], tx[i]);
                sl = A::div(A::sub(sm.L[wsc + i + 1], sm.L[wsc + i]), den);
                zero_dx |= (den == (CarryT)0);
            }
            sm.S[wsc + i] = sl;
        }
        __syncwarp();
        chunk_left = m;
        cbase = pos;
        ctot = tot;
        input_knots += tot;
    };

    auto span_body = [&](auto edge_tag, auto mode_tag, const int sp) {
        constexpr bool EDGE = decltype(edge_tag)::value;
        constexpr int MODE = decltype(mode_tag)::value;
        const bool is_dense = (MODE == 2) ? dense : (MODE == 1);
        const int t0 = sp * SPAN;
        const int tend = t0 + SPAN;                            // first sample after the span
        const bool have_right = !EDGE || tend <= n - 1;
        // ---- many knots: a new chunk's records first (extraction 0 without a scan stage also produces the flag words)
        if (!SCAN && is_dense && chunk_left == 0) {
            if (FIRST) {
                build_chunk_first(sp);
                mc = ld_cg(reinterpret_cast<const uint4 *>(gmask_p() + sp * ITEMS));       // just written (through L2)
            } else {
                build_chunk(sp);
            }
        }
        // ---- early loads: the span's right neighbour (one broadcast load) and the next span's flag words ----
        XT xr = (XT)0;
        uint4 mn = make_uint4(0, 0, 0, 0);
        if (have_right) xr = ld_cg(in_p() + tend);
        if (!SCAN && (!EDGE || sp + 1 < p.spans)) mn = ld_cg(reinterpret_cast<const uint4 *>(gmask_p() + (sp + 1) * ITEMS));
        {
            const int pf = is_dense ? p.pf_dense : p.pf_sparse;
            if (pf > 0 && sp + pf < sp1 && lane < SPAN * (int)sizeof(XT) / 128)
                prefetch_l2(in_p() + t0 + pf * SPAN + lane * (128 / (int)sizeof(XT)));
        }

        const unsigned mw[ITEMS] = {mc.x, mc.y, mc.z, mc.w};
        int wpre[ITEMS];
        wpre[0] = 0;
#pragma unroll
        for (int r = 1; r < ITEMS; ++r) wpre[r] = wpre[r - 1] + __popc(mw[r - 1]);
        const int cnt = SCAN ? 0 : wpre[ITEMS - 1] + __popc(mw[ITEMS - 1]);
        const int fright = (!SCAN && have_right) ? (int)(mn.x & 1u) : 0;

        int ib = 0;                                            // index of the record of the last knot before the span
        if (!SCAN) {
            if (is_dense) {
                --chunk_left;
                ib = wsc + 1 + (pos - cbase);
            } else {
                ib = gbase0 + pos;
            }
        }

        // ---- B, R for the span ------------------------------------------------------------------
        CarryT b[ITEMS];
        if (SCAN) {
#pragma unroll
            for (int r = 0; r < ITEMS; ++r) {
                b[r] = (CarryT)xc[r];
                if (!EDGE || t0 + r * 32 + lane < n) bad |= !isfinite(b[r]);
            }
        } else {
#pragma unroll
            for (int r = 0; r < ITEMS; ++r) {
                const int j = ib + wpre[r] + __popc(mw[r] & le_mask);
                const CarryT xv = (CarryT)xc[r];
                b[r] = A::add(sm.L[j], A::mul(sm.S[j], A::sub(xv, sm.X[j])));     // ITD.py:115-117
                if (EDGE && t0 + r * 32 + lane >= n - 1) b[r] = (CarryT)0;        // ITD.py:112 (and the padding lanes)
            }
            OutT *rot = rot_p() + t0 + lane;
            if (PROBE) {
                // candidate trend row: the input of this extraction (ITD.py:410-411); kRecomp: its baseline
#pragma unroll
                for (int r = 0; r < ITEMS; ++r)
                    if (!EDGE || t0 + r * 32 + lane < n) __stcs(rot + r * 32, RECOMP ? (OutT)b[r] : (OutT)xc[r]);
            } else {
                CarryT *carry = carry_p() + t0 + lane;
                OutT *bas = BAS ? bas_p() + t0 + lane : nullptr;
                if (!last) {
#pragma unroll
                    for (int r = 0; r < ITEMS; ++r) {
                        if (!EDGE || t0 + r * 32 + lane < n) {
                            __stcs(rot + r * 32, (OutT)A::sub((CarryT)xc[r], b[r]));         // ITD.py:119
                            __stwb(carry + r * 32, b[r]);
                            if (BAS) __stcs(bas + r * 32, (OutT)b[r]);
                        }
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < ITEMS; ++r) {
                        if (!EDGE || t0 + r * 32 + lane < n) {
                            const CarryT rr = A::sub((CarryT)xc[r], b[r]);
                            __stcs(rot + r * 32, (OutT)A::add(rr, b[r]));                    // ITD.py:420
                            __stwb(carry + r * 32, b[r]);
                            if (BAS) __stcs(bas + r * 32, (OutT)0);                          // ITD.py:424
                        }
                    }
                }
            }
        }
        // left neighbour of the region's first span
        if (sp == sp0 && sp0 > 0) {
            const CarryT xl = (CarryT)ld_cg(in_p() + t0 - 1);
            bleft = SCAN ? xl : A::add(sm.L[ib], A::mul(sm.S[ib], A::sub(xl, sm.X[ib])));
        }
        // ---- the next span's samples: in flight during the extrema code below -------------------------
        {
            const XT *nx = in_p() + tend + lane;
#pragma unroll
            for (int r = 0; r < ITEMS; ++r) {
                if (EDGE)
                    xc[r] = (tend + r * 32 + lane < n) ? ld_cg(nx + r * 32) : (XT)0;
                else
                    xc[r] = ld_cg(nx + r * 32);
            }
        }
        CarryT bright = (CarryT)0;
        if (have_right) {
            if (SCAN) {
                bright = (CarryT)xr;
            } else if (!EDGE || tend < n - 1) {                                  // B[n-1] is 0 (ITD.py:112)
                const int j = ib + cnt + fright;
                bright = A::add(sm.L[j], A::mul(sm.S[j], A::sub((CarryT)xr, sm.X[j])));
            }
        }

        // ---- extrema of B (of x for the scan): the next level's flag words and knots -------------
        unsigned fw[ITEMS];
        const int newc = span_extrema<EDGE, ITEMS, CarryT>(b, bleft, bright, lane, t0, n, fw);
        if (!PROBE && lane < ITEMS) {
            unsigned v = fw[0];
#pragma unroll
            for (int r = 1; r < ITEMS; ++r) v = (lane == r) ? fw[r] : v;
            __stwb(nmask_p() + sp * ITEMS + lane, v);
        }
        if (!PROBE && newc) {
            int *ntau = reinterpret_cast<int *>(sm.ptr[kPtrNtau]) + roff + kSweepPre + npos;
#ifdef PYITD_SWEEP_XK_LISTS
            CarryT *nxk = reinterpret_cast<CarryT *>(sm.ptr[kPtrNxk]) + roff + kSweepPre + npos;
#endif
            const unsigned lt_mask = le_mask >> 1;
            int pre = 0;
#pragma unroll
            for (int r = 0; r < ITEMS; ++r) {
                if ((fw[r] >> lane) & 1u) {
                    const int rank = pre + __popc(fw[r] & lt_mask);
                    __stwb(ntau + rank, t0 + r * 32 + lane);
#ifdef PYITD_SWEEP_XK_LISTS
                    __stwb(nxk + rank, b[r]);
#endif
                }
                pre += __popc(fw[r]);
            }
        }
        npos += newc;
        pos += cnt;
        bleft = shfl_idx(b[ITEMS - 1], 31);
        mc = mn;
    };

    // EDGE: the span holds sample 0, or it is not followed by a complete span (the last two spans of the signal)
    const int fast_end = min(sp1, n / SPAN - 1);
    int sp = sp0;
    while (sp < sp1) {
        if (sp == 0 || sp >= fast_end) {
            span_body(std::true_type{}, std::integral_constant<int, 2>{}, sp);
            ++sp;
            continue;
        }
        if (SCAN || PROBE || !dense) {
#pragma unroll 1
            for (; sp < fast_end; ++sp) span_body(std::false_type{}, std::integral_constant<int, 0>{}, sp);
        } else {
#pragma unroll 1
            for (; sp < fast_end; ++sp) span_body(std::false_type{}, std::integral_constant<int, 1>{}, sp);
        }
    }
    region_knots = npos;
}

// ---------------------------------------------------------------------------------------------
// one region of one signal, TWO extractions at once (few knots in both: records from the block's two tables).
// Reads X_e once; writes R_e, R_{e+1} and B_{e+1}; B_e exists only in registers: 32 bytes per sample instead of 48 for two
// separate extractions.  The knots of B_e were PREDICTED (sweep_kernel: the extrema stencil at the knots of X_e); this pass
// runs the stencil on every sample of B_e and reports any flag word that differs from the prediction (bad_pred).
//   B_e[t]     = L_k + s_k (X_e[t] - X_k)           table A at index prefix[warp]  + knots of X_e at or before t
//   B_{e+1}[t] = L'_k + s'_k (B_e[t] - X'_k)        table B at index offB + prefix2[warp] + knots of B_e at or before t
// (ITD.py:115-119 twice); the extrema of B_{e+1} are the knots of extraction e + 2.
// ---------------------------------------------------------------------------------------------
template <typename CarryT, typename OutT, bool BAS>
__device__ __forceinline__ void sweep_region_fused(const SweepParams &p, SweepSmem<CarryT> &sm, const int offB,
                                                   const int warp, const int lane, int &region_knots, bool &bad_pred) {
    using A = Arith<CarryT>;
    constexpr int ITEMS = kSweepItems, SPAN = kSweepSpan;
    const int n = p.n;
    const int sp0 = warp * p.spw;
    const int sp1 = min(sp0 + p.spw, p.spans);
    region_knots = 0;
    if (sp0 >= sp1) return;
    auto in_p = [&]() { return reinterpret_cast<const CarryT *>(sm.ptr[kPtrIn]); };
    auto gmask_p = [&]() { return reinterpret_cast<const unsigned *>(sm.ptr[kPtrGmask]); };
    auto gmask2_p = [&]() { return reinterpret_cast<const unsigned *>(sm.ptr[kPtrGmask2]); };
    const int roff = warp * p.rs;
    const int gA = sm.prefix[warp], gB = offB + sm.prefix2[warp];
    const unsigned le_mask = 0xffffffffu >> (31 - lane);
    auto recA = [&](const int j, const CarryT v) { return A::add(sm.L[j], A::mul(sm.S[j], A::sub(v, sm.X[j]))); };

    int posA = 0, posB = 0, npos = 0;
    CarryT bleft = (CarryT)0, bleftA = (CarryT)0;                    // B_{e+1} and B_e left of the span
    CarryT xc[ITEMS];
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
        const int t = sp0 * SPAN + r * 32 + lane;
        xc[r] = (t < n) ? ld_cg(in_p() + t) : (CarryT)0;
    }
    uint4 mc = ld_cg(reinterpret_cast<const uint4 *>(gmask_p() + sp0 * ITEMS));
    uint4 mc2 = ld_cg(reinterpret_cast<const uint4 *>(gmask2_p() + sp0 * ITEMS));

    auto span_body = [&](auto edge_tag, const int sp) {
        constexpr bool EDGE = decltype(edge_tag)::value;
        const int t0 = sp * SPAN, tend = t0 + SPAN;
        const bool have_right = !EDGE || tend <= n - 1;
        CarryT xr = (CarryT)0;
        uint4 mn = make_uint4(0, 0, 0, 0), mn2 = make_uint4(0, 0, 0, 0);
        if (have_right) xr = ld_cg(in_p() + tend);
        if (!EDGE || sp + 1 < p.spans) {
            mn = ld_cg(reinterpret_cast<const uint4 *>(gmask_p() + (sp + 1) * ITEMS));
            mn2 = ld_cg(reinterpret_cast<const uint4 *>(gmask2_p() + (sp + 1) * ITEMS));
        }
        if (p.pf_fused > 0 && sp + p.pf_fused < sp1 && lane < SPAN * (int)sizeof(CarryT) / 128)
            prefetch_l2(in_p() + t0 + p.pf_fused * SPAN + lane * (128 / (int)sizeof(CarryT)));

        const unsigned mw[ITEMS] = {mc.x, mc.y, mc.z, mc.w}, mw2[ITEMS] = {mc2.x, mc2.y, mc2.z, mc2.w};
        int wpre[ITEMS], wpre2[ITEMS];
        wpre[0] = wpre2[0] = 0;
#pragma unroll
        for (int r = 1; r < ITEMS; ++r) {
            wpre[r] = wpre[r - 1] + __popc(mw[r - 1]);
            wpre2[r] = wpre2[r - 1] + __popc(mw2[r - 1]);
        }
        const int cnt = wpre[ITEMS - 1] + __popc(mw[ITEMS - 1]), cnt2 = wpre2[ITEMS - 1] + __popc(mw2[ITEMS - 1]);
        const int fright = have_right ? (int)(mn.x & 1u) : 0, fright2 = have_right ? (int)(mn2.x & 1u) : 0;
        const int ibA = gA + posA, ibB = gB + posB;

        // ---- extraction e: B_e in registers, R_e to its row
        CarryT b[ITEMS];
        {
            OutT *rot = reinterpret_cast<OutT *>(sm.ptr[kPtrRot]) + t0 + lane;
            OutT *bas = BAS ? reinterpret_cast<OutT *>(sm.ptr[kPtrBas]) + t0 + lane : nullptr;
#pragma unroll
            for (int r = 0; r < ITEMS; ++r) {
                b[r] = recA(ibA + wpre[r] + __popc(mw[r] & le_mask), xc[r]);
                if (EDGE && t0 + r * 32 + lane >= n - 1) b[r] = (CarryT)0;       // ITD.py:112 (and the padding lanes)
                if (!EDGE || t0 + r * 32 + lane < n) {
                    __stcs(rot + r * 32, (OutT)A::sub(xc[r], b[r]));              // ITD.py:119
                    if (BAS) __stcs(bas + r * 32, (OutT)b[r]);
                }
            }
        }
        // left neighbour of the region's first span (both levels), then the next span's samples
        if (sp == sp0 && sp0 > 0) {
            bleftA = recA(ibA, ld_cg(in_p() + t0 - 1));
            bleft = recA(ibB, bleftA);
        }
        {
            const CarryT *nx = in_p() + tend + lane;
#pragma unroll
            for (int r = 0; r < ITEMS; ++r) {
                if (EDGE)
                    xc[r] = (tend + r * 32 + lane < n) ? ld_cg(nx + r * 32) : (CarryT)0;
                else
                    xc[r] = ld_cg(nx + r * 32);
            }
        }
        // ---- extraction e + 1 on B_e: R_{e+1} to its row, B_{e+1} to the carry
        CarryT b2[ITEMS];
        {
            OutT *rot2 = reinterpret_cast<OutT *>(sm.ptr[kPtrRot2]) + t0 + lane;
            OutT *bas2 = BAS ? reinterpret_cast<OutT *>(sm.ptr[kPtrBas2]) + t0 + lane : nullptr;
            CarryT *carry = reinterpret_cast<CarryT *>(sm.ptr[kPtrCarry]) + t0 + lane;
#pragma unroll
            for (int r = 0; r < ITEMS; ++r) {
                b2[r] = recA(ibB + wpre2[r] + __popc(mw2[r] & le_mask), b[r]);
                if (EDGE && t0 + r * 32 + lane >= n - 1) b2[r] = (CarryT)0;
                if (!EDGE || t0 + r * 32 + lane < n) {
                    __stcs(rot2 + r * 32, (OutT)A::sub(b[r], b2[r]));
                    __stwb(carry + r * 32, b2[r]);
                    if (BAS) __stcs(bas2 + r * 32, (OutT)b2[r]);
                }
            }
        }
        CarryT bright = (CarryT)0, brightA = (CarryT)0;
        if (have_right && (!EDGE || tend < n - 1)) {                             // both baselines end with 0 (ITD.py:112)
            brightA = recA(ibA + cnt + fright, xr);
            bright = recA(ibB + cnt2 + fright2, brightA);
        }
        // ---- the check: the extrema of B_e are exactly the predicted knots
        {
            unsigned fa[ITEMS];
            span_extrema<EDGE, ITEMS, CarryT>(b, bleftA, brightA, lane, t0, n, fa);
#pragma unroll
            for (int r = 0; r < ITEMS; ++r) bad_pred |= (fa[r] != mw2[r]);
            bleftA = shfl_idx(b[ITEMS - 1], 31);
        }

        // ---- extrema of B_{e+1}: the flag words and knots of extraction e + 2
        unsigned fw[ITEMS];
        const int newc = span_extrema<EDGE, ITEMS, CarryT>(b2, bleft, bright, lane, t0, n, fw);
        if (lane < ITEMS) {
            unsigned v = fw[0];
#pragma unroll
            for (int r = 1; r < ITEMS; ++r) v = (lane == r) ? fw[r] : v;
            __stwb(reinterpret_cast<unsigned *>(sm.ptr[kPtrNmask]) + sp * ITEMS + lane, v);
        }
        if (newc) {
            int *ntau = reinterpret_cast<int *>(sm.ptr[kPtrNtau]) + roff + kSweepPre + npos;
#ifdef PYITD_SWEEP_XK_LISTS
            CarryT *nxk = reinterpret_cast<CarryT *>(sm.ptr[kPtrNxk]) + roff + kSweepPre + npos;
#endif
            const unsigned lt_mask = le_mask >> 1;
            int pre = 0;
#pragma unroll
            for (int r = 0; r < ITEMS; ++r) {
                if ((fw[r] >> lane) & 1u) {
                    const int rank = pre + __popc(fw[r] & lt_mask);
                    __stwb(ntau + rank, t0 + r * 32 + lane);
#ifdef PYITD_SWEEP_XK_LISTS
                    __stwb(nxk + rank, b2[r]);
#endif
                }
                pre += __popc(fw[r]);
            }
        }
        npos += newc;
        posA += cnt;
        posB += cnt2;
        bleft = shfl_idx(b2[ITEMS - 1], 31);
        mc = mn;
        mc2 = mn2;
    };

    const int fast_end = min(sp1, n / SPAN - 1);
    int sp = sp0;
    while (sp < sp1) {
        if (sp == 0 || sp >= fast_end) {
            span_body(std::true_type{}, sp);
            ++sp;
            continue;
        }
#pragma unroll 1
        for (; sp < fast_end; ++sp) span_body(std::false_type{}, sp);
    }
    region_knots = npos;
}

// zero or copy one output row (the knot-stop trend row, zero tails)
template <typename OutT, typename CarryT>
__device__ __forceinline__ void sweep_fill_row(OutT *dst, const CarryT *src, int n, bool zero) {
    constexpr int U = 8;
    for (int t = threadIdx.x; t < n; t += blockDim.x * U) {
        CarryT a[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int tt = t + u * blockDim.x;
            a[u] = (!zero && tt < n) ? ld_cg(src + tt) : (CarryT)0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int tt = t + u * blockDim.x;
            if (tt < n) __stcs(dst + tt, (OutT)a[u]);
        }
    }
}

// The scan-free extraction 0 (kFirst) is an experiment that lost (profiles/r2/README.md): it is compiled only with
// -DPYITD_SWEEP_WITH_FUSED_SCAN (5000 SASS instructions that the default build does not carry).
#ifdef PYITD_SWEEP_WITH_FUSED_SCAN
__device__ __forceinline__ bool kFusedScan(const SweepParams &p) { return p.fused_scan != 0; }
#else
__device__ __forceinline__ constexpr bool kFusedScan(const SweepParams &) { return false; }
#endif

// ---------------------------------------------------------------------------------------------
// sweep_kernel
// ---------------------------------------------------------------------------------------------
// (an 80-register build, 3 CTAs per SM, was measured stage by stage against this one: no stage gains, profiles/r2/README.md)
template <typename InT, typename CarryT, typename OutT, bool BAS>
__global__ void __launch_bounds__(kSweepWarps * 32, 4) sweep_kernel(const SweepParams p) {
    using A = Arith<CarryT>;
#if PYITD_SWEEP_CAP > 2000
    extern __shared__ __align__(16) unsigned char sweep_smem_raw[];
    SweepSmem<CarryT> &sm = *reinterpret_cast<SweepSmem<CarryT> *>(sweep_smem_raw);
#else
    __shared__ SweepSmem<CarryT> sm;        // static: every shared-memory access is a compile-time offset
#endif
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = p.n, S = p.S;
    const bool depth = p.depth_first != 0;
    const long long n_items = depth ? (long long)S : (long long)(p.stage_last - p.stage_first + 1) * S;
    int d_sig = -1, d_e = 0, d_sel = 0, d_nofuse = 0;          // depth-first cursor: the signal this CTA is working through
    // after an item: the next stage of the same signal, or (signal finished) a new ticket
    auto advance = [&](const int e, const bool stopped) {       // e: the last extraction this item completed
        if (e >= p.stage_last || (stopped && !(p.opts & kOptZeroTail))) d_sig = -1;
        else d_e = e + 1;
    };

    // ---- the block's knot table {X, L, S}[off .. off + Kc + 1], one thread per knot (few knots) -----------------------
    // lists: the signal's region lists; pre: knots before each region; el / ex: L and X at the two end knots
    // list slot of the interior knot of the item's input with global rank g (1 <= g <= K)
    auto slot_of = [&](const int g) -> long long {
        int r = 0;
#pragma unroll
        for (int q = 1; q < kSweepWarps; ++q) r += (sm.prefix[q] < g) ? 1 : 0;
        return (long long)r * p.rs + kSweepPre + (g - 1 - sm.prefix[r]);
    };
    // second half of a table build: L (ITD.py:106-110) and the slopes (ITD.py:116) of entries off .. off + Kc + 1, whose tau
    // words (in S's storage) and X are in place
    auto table_compute = [&](const int off, const int Kc, const CarryT *el) {
        int *taus = reinterpret_cast<int *>(sm.S + off);
        for (int k = tid; k <= Kc + 1; k += blockDim.x) {
            CarryT Lv;
            if (k == 0) {
                Lv = el[0];
            } else if (k == Kc + 1) {
                Lv = el[1];
            } else {
                const CarryT w = A::ratio(taus[k] - taus[k - 1], taus[k + 1] - taus[k - 1]);
                const CarryT d = A::sub(sm.X[off + k + 1], sm.X[off + k - 1]);
                const CarryT qq = A::add(sm.X[off + k - 1], A::mul(w, d));
                Lv = A::add(A::mul((CarryT)0.5, qq), A::mul((CarryT)0.5, sm.X[off + k]));
            }
            sm.L[off + k] = Lv;
        }
        __syncthreads();
        bool zdx = false;
        for (int k = tid; k <= Kc + 1; k += blockDim.x) {
            CarryT sl = (CarryT)0;
            if (k <= Kc) {
                const CarryT den = A::sub(sm.X[off + k + 1], sm.X[off + k]);
                sl = A::div(A::sub(sm.L[off + k + 1], sm.L[off + k]), den);
                zdx |= (den == (CarryT)0);
            }
            sm.S[off + k] = sl;                                    // overwrites tau words: see the barrier above
        }
        if (zdx) sm.zero_dx = 1;
        __syncthreads();
    };
    // (x_at: the item's input at a sample -- the lists hold tau only, X_k is gathered)
    auto build_table = [&](const int Kc, const int *ctau, const CarryT *cxk, const CarryT *el, const CarryT *ex, auto x_at) {
        int *taus = reinterpret_cast<int *>(sm.S);                    // tau lives in S's storage until S is computed
        for (int k = tid; k <= Kc + 1; k += blockDim.x) {
            int tv;
            CarryT xv;
            if (k == 0) {
                tv = 0;
                xv = ex[0];
            } else if (k == Kc + 1) {
                tv = n - 1;
                xv = ex[1];
            } else {
                const long long sl = slot_of(k);
                tv = ld_cg(ctau + sl);
#ifdef PYITD_SWEEP_XK_LISTS
                xv = ld_cg(cxk + sl);
#else
                xv = x_at(tv);
#endif
            }
            taus[k] = tv;
            sm.X[k] = xv;
        }
        __syncthreads();
        table_compute(0, Kc, el);
    };

    // ---- fused pair, step 1: the knots of B_e WITHOUT a pass over the signal ----------------------------------------
    // Between two knots of X_e the baseline B_e = L_k + s_k (X_e - X_k) is a monotone function of a monotone stretch of
    // X_e (rounding is monotone), so an extremum of B_e away from the knots of X_e needs two equal neighbouring values of
    // B_e (a flat step) or a sample next to a knot that rounding pushed past it: rare on real data.  So: run the 3-point
    // stencil (ITD.py:44-59) on B_e AT the knots of X_e only -- one thread per knot, B_e at tau_k - 1, tau_k, tau_k + 1 from
    // the table and two gathered samples -- and take the flagged knots as the knots of B_e: their compacted list becomes
    // the second table (at K + 2 in the block's arrays), their flag bits go to the CTA's scratch words.  The fused pass
    // then runs the stencil on EVERY sample of B_e (which it has in registers anyway) and compares: if any flag word
    // differs, the item falls back to a plain extraction e and the signal stops trying pairs.  A wrong prediction costs
    // time, never a wrong bit.
    // Returns the predicted knot count of B_e (a LOWER bound of the true count: every predicted knot is one); go = the
    // second table is in place and a fused pass is possible and worth it (only asked for with want_pair).
    // The count also steers extractions that are not fused: a predicted count below min_extrema means that this
    // extraction is almost certainly the discarded last one (ITD.py:404-411), which is then probed first.
    auto predict_pair = [&](const int K, const int *ctau, const CarryT *x_in, const bool want_pair, bool &go) -> int {
        unsigned *mm = p.mid_mask + (long long)blockIdx.x * p.mstride;
        for (int w = tid; w < (int)p.mstride; w += blockDim.x) __stcg(mm + w, 0u);
        const int offB = K + 2;
        int *tausB = reinterpret_cast<int *>(sm.S + offB);
        auto recA = [&](const int j, const CarryT v) { return A::add(sm.L[j], A::mul(sm.S[j], A::sub(v, sm.X[j]))); };
        __syncthreads();                                               // the scratch words are zero before any bit is set
        // candidates: the knots 1 .. K of X_e, and sample n-2 (B_e[n-1] is forced to 0, ITD.py:112, so n-2 can be an extremum
        // of B_e without being one of X_e -- it is one in every tenth extraction of the benchmark).  Warp w takes the w-th
        // run of candidates, 32 at a time; nothing is synchronised until every load of every step has been issued.
        constexpr int PSTEPS = 7;                                      // 8 warps x 7 x 32 >= kSweepFuseMaxLimit + 1
        const int kw = (K + 1 + kSweepWarps - 1) / kSweepWarps;
        const int kbeg = 1 + warp * kw;
        int tks[PSTEPS];
        CarryT bhs[PSTEPS];
        unsigned bals[PSTEPS];
        int wtot = 0;
#pragma unroll
        for (int st = 0; st < PSTEPS; ++st) {
            const int q = st * 32 + lane, k = kbeg + q;
            bool f = false;
            int tk = 0;
            CarryT bh = (CarryT)0;
            if (q < kw && k <= K) {
                tk = ld_cg(ctau + slot_of(k));
                const int tn = (k + 1 <= K) ? ld_cg(ctau + slot_of(k + 1)) : n - 1;
                const CarryT xp = ld_cg(x_in + tk - 1), xn = ld_cg(x_in + tk + 1);
                const CarryT bp = recA(k - 1, xp);                     // sample tau_k - 1 lies in segment k - 1
                bh = recA(k, sm.X[k]);
                const CarryT bn = (tk + 1 >= n - 1) ? (CarryT)0 : recA((tk + 1 == tn) ? k + 1 : k, xn);      // ITD.py:112
                f = is_knot(bp, bh, bn);
            } else if (q < kw && k == K + 1) {
                const int tK = ld_cg(ctau + slot_of(K));              // (K >= 1)
                if (tK != n - 2) {
                    tk = n - 2;
                    const CarryT bp = recA(K, ld_cg(x_in + n - 3));
                    bh = recA(K, sm.xe[2]);
                    f = is_knot(bp, bh, (CarryT)0);
                }
            }
            bals[st] = __ballot_sync(0xffffffffu, f);
            tks[st] = tk;
            bhs[st] = bh;
            wtot += __popc(bals[st]);
        }
        if (lane == 0) sm.cnt[warp] = wtot;
        __syncthreads();
        int base = 0, K2 = 0;
#pragma unroll
        for (int w = 0; w < kSweepWarps; ++w) {
            const int cw = sm.cnt[w];
            K2 += cw;
            base += (w < warp) ? cw : 0;
        }
#pragma unroll
        for (int st = 0; st < PSTEPS; ++st) {
            if ((bals[st] >> lane) & 1u) {
                const int rank = base + __popc(bals[st] & ((1u << lane) - 1u));
                if (offB + rank + 2 < SweepSmem<CarryT>::kCap) {
                    tausB[1 + rank] = tks[st];
                    sm.X[offB + 1 + rank] = bhs[st];
                }
                atomicOr(mm + (tks[st] >> 5), 1u << (tks[st] & 31));
            }
            base += __popc(bals[st]);
        }
        go = want_pair && (K2 >= p.fuse_min_b) && (K2 >= p.min_extrema) && (K + K2 + 4 <= SweepSmem<CarryT>::kCap);
        if (!go) {
            __syncthreads();                                           // (sm.cnt is reused by the caller)
            return K2;
        }
        if (tid == 0) {
            // the end knots of B_e and ITD.py:100-102 on it: B_e at samples 0, 1, n-2 (n-1 holds 0, ITD.py:112)
            const int t1 = ld_cg(ctau + slot_of(1));                  // (K >= 1)
            const CarryT b0 = recA(0, sm.xe[0]), b1 = recA((t1 == 1) ? 1 : 0, sm.xe[1]);
            const CarryT bz = recA(K, sm.xe[2]);
            tausB[0] = 0;
            sm.X[offB] = b0;
            tausB[K2 + 1] = n - 1;
            sm.X[offB + K2 + 1] = (CarryT)0;
            sm.endl2[0] = mean2<CarryT>(b0, b1);
            sm.endl2[1] = mean2<CarryT>(bz, (CarryT)0);
        }
        __syncthreads();
        if (tid <= kSweepWarps) {                                      // knots of B_e before each region
            const int tstart = (tid == kSweepWarps) ? n : tid * p.spw * kSweepSpan;
            int lo = 1, hi = K2 + 1;                                   // first entry in [1, K2] with tau >= tstart
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (tausB[mid] < tstart) lo = mid + 1;
                else hi = mid;
            }
            sm.prefix2[tid] = lo - 1;
        }
        __syncthreads();
        table_compute(offB, K2, sm.endl2);
        return K2;
    };

    for (;;) {
        __syncthreads();                                       // everyone is done with the previous item's shared state
        int e, sig;
        if (!depth || d_sig < 0) {
            if (tid == 0) sm.ticket = atomicAdd(p.ticket, 1);
            __syncthreads();
            const long long t = sm.ticket;
            if (t >= n_items) break;
            if (depth) {
                d_sig = (int)t;
                d_e = p.stage_first;
            }
            e = p.stage_first + (int)(t / S);                  // -1: scan
            sig = (int)(t % S);
        }
        if (depth) {
            e = d_e;
            sig = d_sig;
        }
        if (p.stage_ns && tid == 0) sm.t_start = global_ns();

        // ---- wait for the previous stage of this signal ----------------------------------------
        bool already = false;
        if (e >= 0 && !(e == 0 && kFusedScan(p)) && (!depth || e == p.stage_first)) {
            if (tid == 0) {
                int dn;
                while (((dn = ld_acquire(p.done + sig)) & kSweepDoneAll) < e + 1) __nanosleep(200);
                __threadfence();
                sm.dn = dn;
            }
            __syncthreads();
            // extraction e was the second of a fused pair (e - 1, e), or the signal has stopped
            already = ((sm.dn & kSweepDoneAll) >= e + 2);
            d_sel = (sm.dn >> 30) & 1;
            d_nofuse = (sm.dn >> 28) & 1;
        } else if (e <= 0) {
            d_sel = (e < 0) ? 1 : 0;                           // the scan writes tab[0]
            d_nofuse = 0;
        }
        const int se = (e >= 0) ? ld_cg(p.stop_e + sig) : kStopOpen;
        if (e > se) {
            // the signal stopped at extraction se: rows beyond its last one are zero-filled on request
            if (p.opts & kOptZeroTail) {
                OutT *rot = reinterpret_cast<OutT *>(p.rot) + (long long)sig * p.out_sig_stride + (long long)e * n;
                sweep_fill_row<OutT, CarryT>(rot, nullptr, n, true);
                if (BAS) {
                    OutT *bas = reinterpret_cast<OutT *>(p.bas) + (long long)sig * p.out_sig_stride + (long long)e * n;
                    sweep_fill_row<OutT, CarryT>(bas, nullptr, n, true);
                }
            }
            advance(e, true);
            continue;
        }
        if (already) {
            advance(e, se <= e);
            continue;
        }

        // ---- per-item setup: base pointers, region prefix, end values ------------------------------
        // d_sel: which table holds the knots of this item's input (see SweepParams): it travels with done[] (bit 30) in
        // stage-major order and in a register in signal-major order.  Everything the item needs later is parked in shared
        // memory: nothing but e, sig and K stays in registers across the span loops.
        if (tid == 64) {
            const int tsel = d_sel;
            sm.it_e = e;
            sm.it_sig = sig;
            sm.it_sel = d_sel;
            sm.it_nofuse = d_nofuse;
            sm.it_de = d_e;
            sm.it_dsig = d_sig;
            const long long koff = (long long)sig * p.kstride, moff = (long long)sig * p.mstride;
            const SweepTable &cur = p.tab[tsel], &nxt = p.tab[tsel ^ 1];
            const CarryT *x_in = reinterpret_cast<const CarryT *>(p.carry[tsel ^ 1]) + (long long)sig * n;     // X_e, e >= 1
            const long long row = (long long)sig * p.out_sig_stride + (long long)e * n;
            sm.ptr[kPtrIn] = (e <= 0) ? (void *)(reinterpret_cast<const InT *>(p.x) + (long long)sig * n) : (void *)x_in;
            sm.ptr[kPtrRot] = reinterpret_cast<OutT *>(p.rot) + row;
            sm.ptr[kPtrBas] = BAS ? reinterpret_cast<OutT *>(p.bas) + row : nullptr;
            sm.ptr[kPtrCarry] = reinterpret_cast<CarryT *>(p.carry[tsel]) + (long long)sig * n;
            sm.ptr[kPtrGmask] = cur.mask + moff;
            sm.ptr[kPtrNmask] = nxt.mask + moff;
            sm.ptr[kPtrCtau] = cur.tau + koff;
            sm.ptr[kPtrCxk] = reinterpret_cast<CarryT *>(cur.xk) + koff;
            sm.ptr[kPtrNtau] = nxt.tau + koff;
            sm.ptr[kPtrNxk] = reinterpret_cast<CarryT *>(nxt.xk) + koff;
        }
        int K = 0;
        bool dense = false, can_fuse = false;
        if (e < 0) __syncthreads();
        if (e >= 0) {
            const bool first_fused = (e == 0) && kFusedScan(p);
            const SweepTable &cur = p.tab[d_sel];
            const long long koff = (long long)sig * p.kstride;
            if (tid == 0) {
                int run = 0;
                for (int r = 0; r < kSweepWarps; ++r) {
                    sm.prefix[r] = run;
                    // (extraction 0 without a scan stage: the knots are not known yet; it always runs in chunk mode)
                    run += first_fused ? kSweepCap : ld_cg(cur.rcount + (long long)sig * kSweepWarps + r);
                }
                sm.prefix[kSweepWarps] = run;
                sm.zero_dx = 0;
            }
            if (tid == 32) {
                // ITD.py:100-102 and the values at the two end knots
                CarryT a0, a1, z0, z1;
                if (e == 0) {
                    const InT *xi = reinterpret_cast<const InT *>(p.x) + (long long)sig * n;
                    a0 = (CarryT)xi[0]; a1 = (CarryT)xi[1]; z0 = (CarryT)xi[n - 2]; z1 = (CarryT)xi[n - 1];
                } else {
                    const CarryT *x_in = reinterpret_cast<const CarryT *>(p.carry[d_sel ^ 1]) + (long long)sig * n;
                    a0 = ld_cg(x_in); a1 = ld_cg(x_in + 1); z0 = ld_cg(x_in + n - 2); z1 = ld_cg(x_in + n - 1);
                }
                sm.endl[0] = mean2<CarryT>(a0, a1);
                sm.endl[1] = mean2<CarryT>(z0, z1);
                sm.endx[0] = a0;
                sm.endx[1] = z1;
                sm.xe[0] = a0, sm.xe[1] = a1, sm.xe[2] = z0, sm.xe[3] = z1;
            }
            __syncthreads();
            K = sm.prefix[kSweepWarps];
            dense = (K + 2 > SweepSmem<CarryT>::kCap);
#ifdef PYITD_SWEEP_NO_FUSE_CODE
            can_fuse = false;
#else
            can_fuse = p.fuse && !d_nofuse && !dense && e >= 1 && K >= p.fuse_min_a && K + 2 <= p.fuse_max_a &&
                       e + 1 < p.emax && e + 1 <= p.stage_last && (int)blockIdx.x < p.mid_ctas;
#endif
            const int *ctau = cur.tau + koff;
            const CarryT *cxk = reinterpret_cast<const CarryT *>(cur.xk) + koff;
            if (dense) {
                // ---- halo slots of every region list: the two knots before and the three after the region ----
                if (!first_fused && tid < kSweepWarps * (kSweepPre + kSweepPost)) {
                    const int r = tid / (kSweepPre + kSweepPost), h = tid % (kSweepPre + kSweepPost);
                    const int c = sm.prefix[r + 1] - sm.prefix[r];
                    const int j = (h < kSweepPre) ? h - kSweepPre : c + (h - kSweepPre);       // local index in region r
                    const int g = 1 + sm.prefix[r] + j;
                    int tv = 0;
                    CarryT xv = (CarryT)0;
                    if (g == 0) {
                        tv = 0;
                        xv = sm.endx[0];
                    } else if (g == K + 1) {
                        tv = n - 1;
                        xv = sm.endx[1];
                    } else if (g >= 1 && g <= K) {
                        const long long sl = slot_of(g);
                        tv = ld_cg(ctau + sl);
#ifdef PYITD_SWEEP_XK_LISTS
                        xv = ld_cg(cxk + sl);
#endif
                    }
                    const long long dst = (long long)r * p.rs + kSweepPre + j;
                    const_cast<int *>(ctau)[dst] = tv;
#ifdef PYITD_SWEEP_XK_LISTS
                    const_cast<CarryT *>(cxk)[dst] = xv;
#else
                    (void)xv;
#endif
                }
                __syncthreads();
            } else {
                if (e == 0) {
                    const InT *xi = reinterpret_cast<const InT *>(p.x) + (long long)sig * n;
                    build_table(K, ctau, cxk, sm.endl, sm.endx, [&](const int t) { return (CarryT)xi[t]; });
                } else {
                    const CarryT *xi = reinterpret_cast<const CarryT *>(p.carry[d_sel ^ 1]) + (long long)sig * n;
                    build_table(K, ctau, cxk, sm.endl, sm.endx, [&](const int t) { return ld_cg(xi + t); });
                }
            }
        }

        // ---- stream the regions (no block barrier inside) ----------------------------------------
        int region_knots = 0;
        bool zero_dx = false, bad = false;
        const bool last0 = (e == p.emax);
        // ---- the knots of B_e, predicted (few knots, e >= 1): decides between probe, fused pair and plain extraction
        int K2 = 0;
        bool go = false, predicted = false;
        if (p.predict && !d_nofuse && !dense && e >= 1 && !last0 && K >= 1 && K <= kSweepFuseMaxLimit &&
            (int)blockIdx.x < p.mid_ctas) {
            const CarryT *x_in = reinterpret_cast<const CarryT *>(p.carry[d_sel ^ 1]) + (long long)sig * n;
            K2 = predict_pair(K, p.tab[d_sel].tau + (long long)sig * p.kstride, x_in, can_fuse, go);
            predicted = true;
            if (can_fuse && !go && tid == 0) atomicAdd(p.stats + 1, 1);
        }
        // an extraction whose baseline is predicted to have fewer than min_extrema extrema (without a prediction: one with at
        // most kSweepProbeKnots knots) is probably the discarded last one: probe it first
        bool probed_stop = false;
        if (e >= 1 && !last0 && !go && (predicted ? (K2 < p.min_extrema) : (K <= kSweepProbeKnots))) {
            int unused = 0;
            sweep_region<CarryT, CarryT, OutT, kProbe, BAS>(p, sm, false, false, K, warp, lane, region_knots, unused, zero_dx, bad);
            if (lane == 0) sm.cnt[warp] = region_knots;
            __syncthreads();
            int kp = 0;
#pragma unroll
            for (int r = 0; r < kSweepWarps; ++r) kp += sm.cnt[r];
            probed_stop = (kp < p.min_extrema);
            __syncthreads();
        }
        // ---- a fused pair: extractions e and e + 1 in ONE pass over X_e; B_e is never stored ---------------------------
        // The knots of B_e are predicted from the table (predict_pair); the pass reads X_e and writes R_e, R_{e+1}, B_{e+1}
        // (32 bytes per sample instead of 48 for ITD.py:79-121 twice) and checks the prediction on every sample.  The stop
        // test of extraction e (ITD.py:404) is the predicted count: a pair is only tried when it says "go on".
        int fused = 0;
        bool pair_failed = false;
        if (go) {
            {
                if (tid == 64) {
                    const long long row2 = (long long)sig * p.out_sig_stride + (long long)(e + 1) * n;
                    sm.ptr[kPtrRot2] = reinterpret_cast<OutT *>(p.rot) + row2;
                    sm.ptr[kPtrBas2] = BAS ? reinterpret_cast<OutT *>(p.bas) + row2 : nullptr;
                    sm.ptr[kPtrGmask2] = p.mid_mask + (long long)blockIdx.x * p.mstride;
                    sm.pred_bad = 0;
                }
                __syncthreads();
                bool bad_pred = false;
                sweep_region_fused<CarryT, OutT, BAS>(p, sm, K + 2, warp, lane, region_knots, bad_pred);
                if (__any_sync(0xffffffffu, bad_pred) && lane == 0) sm.pred_bad = 1;
                if (lane == 0) sm.cnt[warp] = region_knots;
                __syncthreads();
                e = sm.it_e, sig = sm.it_sig, d_sel = sm.it_sel, d_e = sm.it_de, d_sig = sm.it_dsig, K = sm.prefix[kSweepWarps];
                d_nofuse = sm.it_nofuse;
                pair_failed = (sm.pred_bad != 0);
                if (tid == 0) atomicAdd(p.stats + (pair_failed ? 2 : 0), 1);
                if (pair_failed) {
                    // B_e has an extremum the prediction did not see (a flat step, a tie next to a knot): everything the pass
                    // wrote is overwritten by the plain extraction e below (row e, the carry, the next knot lists) or by
                    // whatever later fills row e + 1; this signal does not try pairs again
                    d_nofuse = 1;
                    K2 = 0;
                    __syncthreads();                                   // (sm.cnt is written again below)
                } else {
                    fused = 1;
                }
            }
        }
        if (tid == 0) {
            sm.it_flags = fused | (probed_stop ? 4 : 0);
            sm.it_k2 = K2;
            if (pair_failed) sm.it_nofuse = 1;                         // (behind the barrier of the failed path: nobody reads it now)
        }
        int knots_in = 0;
        if (probed_stop || fused) {
            // probed_stop: row e already holds the trend row; the region counts of the probe are the ones to report
        } else if (e < 0) {
            sweep_region<InT, CarryT, OutT, kScan, BAS>(p, sm, false, false, 0, warp, lane, region_knots, knots_in, zero_dx, bad);
        } else if (e == 0 && kFusedScan(p)) {
            sweep_region<InT, CarryT, OutT, kFirst, BAS>(p, sm, true, last0, K, warp, lane, region_knots, knots_in, zero_dx, bad);
            if (lane == 0) sm.knots_in[warp] = knots_in;
        } else if (std::is_same<InT, CarryT>::value || e > 0) {
            sweep_region<CarryT, CarryT, OutT, kLevel, BAS>(p, sm, dense, last0, K, warp, lane, region_knots, knots_in, zero_dx, bad);
        } else {
            sweep_region<InT, CarryT, OutT, kLevel, BAS>(p, sm, dense, last0, K, warp, lane, region_knots, knots_in, zero_dx, bad);
        }
        if (lane == 0) sm.cnt[warp] = region_knots;
        if (__any_sync(0xffffffffu, zero_dx) && lane == 0) sm.zero_dx = 1;
        if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(p.status + sig, kStNonFinite);
        __syncthreads();

        // ---- end of the item: region counts, stop rule, trend row ---------------------------------
        e = sm.it_e, sig = sm.it_sig, d_sel = sm.it_sel, d_e = sm.it_de, d_sig = sm.it_dsig, K = sm.prefix[kSweepWarps];
        fused = sm.it_flags & 1, probed_stop = (sm.it_flags & 4) != 0, K2 = sm.it_k2, d_nofuse = sm.it_nofuse;
        const bool last = (e == p.emax);
        const int ee = e + fused;                                      // the last extraction this item completed
        int Kn = 0;
#pragma unroll
        for (int r = 0; r < kSweepWarps; ++r) Kn += sm.cnt[r];
        if (tid < kSweepWarps) p.tab[d_sel ^ 1].rcount[(long long)sig * kSweepWarps + tid] = sm.cnt[tid];
        bool stop_knots = false;
        if (e < 0) {
            if (tid == 0 && p.input_knots) p.input_knots[sig] = Kn;
        } else {
            stop_knots = (Kn < p.min_extrema);                         // ITD.py:404
            if (tid == 0 && e == 0 && kFusedScan(p) && p.input_knots) {
                int kin = 0;
                for (int r = 0; r < kSweepWarps; ++r) kin += sm.knots_in[r];
                p.input_knots[sig] = kin;
            }
            if (tid == 0) {
                if (sm.zero_dx) atomicOr(p.status + sig, kStZeroDx);
                if (fused) p.knot_counts[(long long)sig * p.rows + e] = K2;
                p.knot_counts[(long long)sig * p.rows + ee] = Kn;      // what ITD.py:403 prints
                if (stop_knots || last) {                              // ITD.py:404 / :418
                    p.stop_kind[sig] = stop_knots ? kStopKnots : kStopIter;
                    p.n_rows[sig] = ee + 1;
                    p.stop_e[sig] = ee;
                }
            }
            if (stop_knots && !probed_stop && !fused) {
                // the discarded extraction wrote R_e into row e; the reference returns
                // baselines[e-1] there, i.e. the INPUT of this extraction (zeros when e == 0)  (ITD.py:410-411)
                OutT *rot = reinterpret_cast<OutT *>(p.rot) + (long long)sig * p.out_sig_stride + (long long)e * n;
                const CarryT *x_in = reinterpret_cast<const CarryT *>(p.carry[d_sel ^ 1]) + (long long)sig * n;
                sweep_fill_row<OutT, CarryT>(rot, (e == 0) ? nullptr : x_in, n, e == 0);
            }
            if (stop_knots && fused) {
                // the second extraction of the pair is the discarded one: row e + 1 is its input B_e, which was never
                // stored -- evaluate it again from X_e and the first table (rare: one extra read and write of the signal)
                if (tid == 64) sm.ptr[kPtrRot] = sm.ptr[kPtrRot2];
                __syncthreads();
                int u0 = 0, u1 = 0;
                bool z = false, b = false;
                sweep_region<CarryT, CarryT, OutT, kRecomp, BAS>(p, sm, false, false, K, warp, lane, u0, u1, z, b);
            }
            if (stop_knots && BAS && (p.opts & kOptZeroTail)) {
                __syncthreads();
                OutT *bas = reinterpret_cast<OutT *>(p.bas) + (long long)sig * p.out_sig_stride + (long long)ee * n;
                sweep_fill_row<OutT, CarryT>(bas, nullptr, n, true);
            }
        }
        __syncthreads();                                               // every thread's global writes are issued
        if (tid == 0) {
            __threadfence();
            // a stopped signal lets every later stage of it through at once (they only zero-fill on request)
            st_release(p.done + sig,
                       ((e >= 0 && (stop_knots || last)) ? kSweepDoneAll : ee + 2) | ((d_sel ^ 1) << 30) | (d_nofuse << 28));
            if (p.stage_ns) atomicAdd(p.stage_ns + (e + 1), global_ns() - sm.t_start);
        }
        d_sel ^= 1;
        advance(ee, e >= 0 && (stop_knots || last));
    }
}

}  // namespace pyitd
