// itd_coop.cuh -- a handful of signals (the reference's own use: ITD().itd(x) on ONE signal, ITD.py:500-503): the whole
// decomposition in ONE cooperative launch with the signal ON CHIP.
//
// A group of CTAs owns one signal; CTA c of the group keeps samples [c C, (c+1) C) of the current input X_e (plus one
// halo sample on each side) in shared memory, in the carry type, for the whole decomposition.  Per extraction
// (ITD.py:79-121) a CTA needs from the rest of the signal only
//   * the two knots before and the three knots after its chunk (ITD.py:106-110 looks one knot to each side, the
//     segment of the right halo sample one further),
//   * the knot count of the whole signal (the stop test, ITD.py:400-404),
//   * X_e at the four end samples (ITD.py:100-102),
// so every CTA publishes a 100-byte summary of its chunk (knot count, first three / last two knots, end samples) and the
// group meets at ONE barrier per extraction (a counter in L2: arrive with a release, spin on an acquire).  Everything
// else -- the extrema stencil, the compaction of the chunk's knots, L_k and the slopes (one thread per knot), the
// evaluation of B and R -- happens in shared memory; global memory sees the input once and every output row once.
// The B of the halo samples is computed by both neighbours (same operands, same operations: same bits), so no sample
// value is ever exchanged.
//
// 65 536 samples on 128 CTAs: ~12 extractions x (one barrier + a few microseconds) instead of the ~30 launches of the
// look-back kernels.  Same arithmetic and operation order as every other path: bit-identical to the reference in fp64.
#pragma once

#include "itd_kernels.cuh"

namespace pyitd {

constexpr int kCoopThreads = 256;
constexpr int kCoopWarps = kCoopThreads / 32;
constexpr int kCoopPre = 2, kCoopPost = 3;            // knots taken from before / after the chunk
constexpr int kCoopMaxChunk = 4096;                   // samples per CTA (shared memory: ~46 bytes per sample)

template <typename CarryT>
struct CoopSummary {
    int cnt;                 // knots of the chunk
    int tfirst[3];           // its first three knots ...
    int tlast[2];            // ... and its last two (tlast[1] is the last one)
    int pad[2];
    CarryT xfirst[3], xlast[2];
    CarryT edge[4];          // chunk 0 fills [0..1] = X[0], X[1]; the last chunk fills [2..3] = X[n-2], X[n-1]
};

struct CoopParams {
    const void *x;           // [S, n] input type
    void *rot, *bas;         // [S, rows, n] output type; bas may be null
    long long out_sig_stride;
    int *n_rows, *knot_counts, *input_knots, *stop_kind, *status;
    int *bar;                // [2 * ngroups]: arrivals, exits (zero between launches)
    void *sum;               // [2][ngroups][gsz] CoopSummary<CarryT>
    int S, n;
    int C, gsz, ngroups;     // chunk length, CTAs per signal, signals in flight
    int emax, rows, min_extrema;
    unsigned opts;
};

template <typename CarryT>
__host__ __device__ inline size_t coop_smem_bytes(int C) {
    return (size_t)(2 * (C + 2) + 3 * (C + 8)) * sizeof(CarryT) + (size_t)(C + 8) * sizeof(int) +
           (size_t)(C + 8) * sizeof(unsigned short) + 256;
}

__device__ __forceinline__ int coop_ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <typename InT, typename CarryT, typename OutT, bool BAS>
__global__ void __launch_bounds__(kCoopThreads) coop_kernel(const CoopParams p) {
    using A = Arith<CarryT>;
    using Sum = CoopSummary<CarryT>;
    extern __shared__ __align__(16) unsigned char coop_raw[];
    const int C = p.C, n = p.n, gsz = p.gsz;
    CarryT *buf0 = reinterpret_cast<CarryT *>(coop_raw);          // [C + 2]: element 0 is the left halo
    CarryT *buf1 = buf0 + (C + 2);
    CarryT *tx = buf1 + (C + 2);                                  // knot table: kCoopPre before, the chunk's, kCoopPost after
    CarryT *tL = tx + (C + 8);
    CarryT *tS = tL + (C + 8);
    int *ttau = reinterpret_cast<int *>(tS + (C + 8));
    unsigned short *seg = reinterpret_cast<unsigned short *>(ttau + (C + 8));     // knots of the chunk at or before a sample
    __shared__ int wcnt[2][kCoopWarps];
    __shared__ int s_K, s_flags;
    __shared__ CarryT s_endl[2], s_endx[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int group = blockIdx.x / gsz, c = blockIdx.x % gsz;
    const int t0 = c * C;
    const int clen = min(C, n - t0);                               // >= 1 by construction of gsz
    const int t1 = t0 + clen;
    int *bar = p.bar + 2 * group;
    int bar_target = 0;
    int pub = 0;                                                   // summaries published so far: slot pub & 1 (see below)
    const unsigned lt_mask = (1u << lane) - 1u;

    // the group barrier: every CTA's writes before it are visible to every CTA's reads after it
    auto group_barrier = [&]() {
        __syncthreads();
        bar_target += gsz;
        if (tid == 0) {
            __threadfence();
            atomicAdd(bar, 1);
            unsigned spins = 0;
            while (coop_ld_acquire(bar) < bar_target) {
                if (++spins > (1u << 27)) __trap();                // seconds: a lost CTA must not hang the device
            }
            __threadfence();
        }
        __syncthreads();
    };

    // knots of the chunk of `cur` (cur[-1] and cur[clen] are the halo samples): compacted into the table at kCoopPre,
    // seg[j] = knots at or before local sample j.  ITD.py:44-59 on x and -x, 1 <= t <= n-2 (ITD.py:70-73)
    auto find_knots = [&](const CarryT *cur) -> int {
        int base = 0, par = 0;
        for (int j0 = 0; j0 < clen; j0 += kCoopThreads) {
            const int j = j0 + tid, t = t0 + j;
            const bool f = (j < clen) && t >= 1 && t <= n - 2 && is_knot(cur[j - 1], cur[j], cur[j + 1]);
            const unsigned bal = __ballot_sync(0xffffffffu, f);
            if (lane == 0) wcnt[par][warp] = __popc(bal);
            __syncthreads();
            int pre = 0, tot = 0;
#pragma unroll
            for (int w = 0; w < kCoopWarps; ++w) {
                const int cw = wcnt[par][w];
                tot += cw;
                pre += (w < warp) ? cw : 0;
            }
            const int rank = base + pre + __popc(bal & lt_mask);
            if (f) {
                ttau[kCoopPre + rank] = t;
                tx[kCoopPre + rank] = cur[j];
            }
            if (j < clen) seg[j] = (unsigned short)(rank + (f ? 1 : 0));
            base += tot;
            par ^= 1;
        }
        __syncthreads();
        return base;
    };

    for (int sig = group; sig < p.S; sig += p.ngroups) {
        Sum *sums = reinterpret_cast<Sum *>(p.sum);
        OutT *rot = reinterpret_cast<OutT *>(p.rot) + (long long)sig * p.out_sig_stride;
        OutT *bas = BAS ? reinterpret_cast<OutT *>(p.bas) + (long long)sig * p.out_sig_stride : nullptr;
        CarryT *cur = buf0 + 1, *oth = buf1 + 1;

        // ---- the input: chunk + halos into shared memory, in the carry type
        bool bad = false;
        {
            const InT *xi = reinterpret_cast<const InT *>(p.x) + (long long)sig * n;
            for (int j = tid - 1; j <= clen; j += kCoopThreads) {
                const int t = t0 + j;
                CarryT v = (CarryT)0;
                if (t >= 0 && t < n) {
                    v = (CarryT)xi[t];
                    if (j >= 0 && j < clen) bad |= !isfinite(v);
                }
                cur[j] = v;
            }
        }
        if (c == 0 && tid == 0) {
            p.status[sig] = 0;
            p.n_rows[sig] = 0;
            if (p.stop_kind) p.stop_kind[sig] = 0;
            for (int r = 0; r < p.rows; ++r) p.knot_counts[(long long)sig * p.rows + r] = 0;
        }
        const int anybad = __syncthreads_or(bad ? 1 : 0);
        if (tid == 0) s_flags = anybad ? kStNonFinite : 0;
        __syncthreads();
        int lk = find_knots(cur);
        bool stopped_knots = false;
        int n_rows = 0;

        for (int e = 0;; ++e) {
            // ---- publish the chunk's summary of X_e, meet the group.  Two slots, used alternately ACROSS signals: a CTA
            // reads publication k after barrier k and before it arrives at barrier k + 1, and publication k + 2 (the
            // next one into the same slot) is written after barrier k + 1
            const size_t slot = ((size_t)(pub & 1) * p.ngroups + group) * gsz;
            ++pub;
            Sum *mine = sums + slot + c;
            if (tid < 3) {
                __stcg(&mine->tfirst[tid], (tid < lk) ? ttau[kCoopPre + tid] : -1);
                __stcg(&mine->xfirst[tid], (tid < lk) ? tx[kCoopPre + tid] : (CarryT)0);
            } else if (tid < 5) {
                const int q = tid - 3;                              // 0: second to last, 1: last
                const int i = lk - 2 + q;
                __stcg(&mine->tlast[q], (i >= 0) ? ttau[kCoopPre + i] : -1);
                __stcg(&mine->xlast[q], (i >= 0) ? tx[kCoopPre + i] : (CarryT)0);
            } else if (tid == 5) {
                __stcg(&mine->cnt, lk);
            } else if (tid == 6 && c == 0) {
                __stcg(&mine->edge[0], cur[0]);
                __stcg(&mine->edge[1], cur[1 - t0]);                // (t0 == 0)
            } else if (tid == 7 && c == gsz - 1) {
                __stcg(&mine->edge[2], cur[n - 2 - t0]);            // (may be the left halo)
                __stcg(&mine->edge[3], cur[n - 1 - t0]);
            }
            group_barrier();

            // ---- the rest of the signal, as far as this chunk needs it
            const Sum *gs = sums + slot;
            if (warp == 0) {
                int have = 0, pos = c - 1;
                while (have < kCoopPre && pos >= 0) {
                    const int idx = pos - lane;
                    const int cv = (idx >= 0) ? __ldcg(&gs[idx].cnt) : 0;
                    const unsigned ball = __ballot_sync(0xffffffffu, cv > 0);
                    if (!ball) {
                        pos -= 32;
                        continue;
                    }
                    const int first = __ffs(ball) - 1, q = pos - first;
                    const int take = min(__shfl_sync(0xffffffffu, cv, first), kCoopPre - have);
                    if (lane < take) {                              // lane 0: the chunk's last knot, lane 1: the one before
                        ttau[kCoopPre - 1 - have - lane] = __ldcg(&gs[q].tlast[1 - lane]);
                        tx[kCoopPre - 1 - have - lane] = __ldcg(&gs[q].xlast[1 - lane]);
                    }
                    have += take;
                    pos = q - 1;
                }
                if (lane == 0) {
                    if (have < kCoopPre) {                          // the end knot at sample 0 (ITD.py:98)
                        ttau[kCoopPre - 1 - have] = 0;
                        tx[kCoopPre - 1 - have] = __ldcg(&gs[0].edge[0]);
                        ++have;
                    }
                    for (; have < kCoopPre; ++have) ttau[kCoopPre - 1 - have] = -1;
                }
            } else if (warp == 1) {
                int have = 0, pos = c + 1;
                while (have < kCoopPost && pos < gsz) {
                    const int idx = pos + lane;
                    const int cv = (idx < gsz) ? __ldcg(&gs[idx].cnt) : 0;
                    const unsigned ball = __ballot_sync(0xffffffffu, cv > 0);
                    if (!ball) {
                        pos += 32;
                        continue;
                    }
                    const int first = __ffs(ball) - 1, q = pos + first;
                    const int take = min(__shfl_sync(0xffffffffu, cv, first), kCoopPost - have);
                    if (lane < take) {
                        ttau[kCoopPre + lk + have + lane] = __ldcg(&gs[q].tfirst[lane]);
                        tx[kCoopPre + lk + have + lane] = __ldcg(&gs[q].xfirst[lane]);
                    }
                    have += take;
                    pos = q + 1;
                }
                if (lane == 0) {
                    if (have < kCoopPost) {                         // the end knot at sample n-1
                        ttau[kCoopPre + lk + have] = n - 1;
                        tx[kCoopPre + lk + have] = __ldcg(&gs[gsz - 1].edge[3]);
                        ++have;
                    }
                    for (; have < kCoopPost; ++have) ttau[kCoopPre + lk + have] = -1;
                }
            } else if (warp == 2) {
                int k = 0;
                for (int i = lane; i < gsz; i += 32) k += __ldcg(&gs[i].cnt);
                k = __reduce_add_sync(0xffffffffu, k);
                if (lane == 0) {
                    s_K = k;
                    const CarryT a0 = __ldcg(&gs[0].edge[0]), a1 = __ldcg(&gs[0].edge[1]);
                    const CarryT z0 = __ldcg(&gs[gsz - 1].edge[2]), z1 = __ldcg(&gs[gsz - 1].edge[3]);
                    s_endl[0] = mean2<CarryT>(a0, a1);              // ITD.py:101-102
                    s_endl[1] = mean2<CarryT>(z0, z1);
                    s_endx[0] = a0;
                    s_endx[1] = z1;
                }
            }
            __syncthreads();
            const int K = s_K;

            // ---- the stop test of the previous extraction (ITD.py:400-426): K = extrema of B_{e-1}
            if (e == 0) {
                if (c == 0 && tid == 0 && p.input_knots) p.input_knots[sig] = K;
            } else {
                if (c == 0 && tid == 0) p.knot_counts[(long long)sig * p.rows + (e - 1)] = K;       // what ITD.py:403 prints
                if (K < p.min_extrema) {
                    // extraction e-1 is discarded: its row is its INPUT (zeros for e-1 == 0), ITD.py:404-411
                    OutT *row = rot + (long long)(e - 1) * n + t0;
                    for (int j = tid; j < clen; j += kCoopThreads) row[j] = (e - 1 == 0) ? (OutT)0 : (OutT)oth[j];
                    stopped_knots = true;
                    n_rows = e;
                    break;
                }
                if (e - 1 == p.emax) {                              // ITD.py:418-426: the row already holds R + B
                    n_rows = e;
                    break;
                }
            }

            // ---- L_k (ITD.py:106-110) and the slopes (ITD.py:116) of the table, one thread per entry
            const int m = kCoopPre + lk + kCoopPost;
            for (int i = tid; i < m; i += kCoopThreads) {
                const int ti = ttau[i];
                CarryT Lv = (CarryT)0;
                if (ti == 0) {
                    Lv = s_endl[0];
                } else if (ti == n - 1) {
                    Lv = s_endl[1];
                } else if (ti > 0 && i >= 1 && i + 1 < m && ttau[i - 1] >= 0 && ttau[i + 1] >= 0) {
                    const int tl = ttau[i - 1];
                    const CarryT w = A::ratio(ti - tl, ttau[i + 1] - tl);
                    const CarryT xl = tx[i - 1];
                    const CarryT d = A::sub(tx[i + 1], xl);
                    const CarryT qq = A::add(xl, A::mul(w, d));
                    Lv = A::add(A::mul((CarryT)0.5, qq), A::mul((CarryT)0.5, tx[i]));
                }
                tL[i] = Lv;
            }
            __syncthreads();
            bool zdx = false;
            for (int i = tid; i < m; i += kCoopThreads) {
                CarryT sl = (CarryT)0;
                // segments that start at the last knot before the chunk ... at the first knot after it
                if (i >= kCoopPre - 1 && i <= kCoopPre + lk && ttau[i] >= 0 && ttau[i] != n - 1 && ttau[i + 1] >= 0) {
                    const CarryT den = A::sub(tx[i + 1], tx[i]);
                    sl = A::div(A::sub(tL[i + 1], tL[i]), den);
                    zdx |= (den == (CarryT)0);
                }
                tS[i] = sl;
            }
            if (zdx) atomicOr(&s_flags, kStZeroDx);
            __syncthreads();

            // ---- B_e into the other buffer (halo samples included), R_e to its row (ITD.py:112-119)
            const bool last = (e == p.emax);
            {
                OutT *row = rot + (long long)e * n + t0;
                OutT *brow = BAS ? bas + (long long)e * n + t0 : nullptr;
                for (int j = tid; j < clen; j += kCoopThreads) {
                    const int i = kCoopPre - 1 + seg[j];
                    const CarryT xv = cur[j];
                    CarryT b = A::add(tL[i], A::mul(tS[i], A::sub(xv, tx[i])));
                    if (t0 + j >= n - 1) b = (CarryT)0;             // ITD.py:112
                    oth[j] = b;
                    const CarryT rr = A::sub(xv, b);
                    row[j] = last ? (OutT)A::add(rr, b) : (OutT)rr; // ITD.py:420 / :119
                    if (BAS) brow[j] = last ? (OutT)0 : (OutT)b;    // ITD.py:424
                }
                if (tid == 0 && t0 > 0) {
                    const int i = kCoopPre - 1;
                    oth[-1] = A::add(tL[i], A::mul(tS[i], A::sub(cur[-1], tx[i])));
                }
                if (tid == 32 && t1 < n) {
                    const int i = kCoopPre - 1 + lk + ((ttau[kCoopPre + lk] == t1) ? 1 : 0);
                    CarryT b = A::add(tL[i], A::mul(tS[i], A::sub(cur[clen], tx[i])));
                    if (t1 >= n - 1) b = (CarryT)0;
                    oth[clen] = b;
                }
            }
            __syncthreads();
            {
                CarryT *sw = cur;
                cur = oth;
                oth = sw;
            }
            lk = find_knots(cur);
        }

        // ---- end of the signal
        if (c == 0 && tid == 0) {
            p.n_rows[sig] = n_rows;
            if (p.stop_kind) p.stop_kind[sig] = stopped_knots ? kStopKnots : kStopIter;
        }
        __syncthreads();
        if (tid == 0 && s_flags) atomicOr(p.status + sig, s_flags);
        if (p.opts & kOptZeroTail) {
            for (int r = stopped_knots ? n_rows - 1 : n_rows; r < p.rows; ++r) {
                if (r >= n_rows) {
                    OutT *row = rot + (long long)r * n + t0;
                    for (int j = tid; j < clen; j += kCoopThreads) row[j] = (OutT)0;
                }
                if (BAS) {
                    OutT *brow = bas + (long long)r * n + t0;
                    for (int j = tid; j < clen; j += kCoopThreads) brow[j] = (OutT)0;
                }
            }
        }
        __syncthreads();                                            // s_flags, the buffers: before the next signal
    }

    // ---- leave the counters at zero for the next launch: the group's first CTA waits for everyone to be done with them
    if (tid == 0) {
        __threadfence();
        atomicAdd(bar + 1, 1);
        if (c == 0) {
            unsigned spins = 0;
            while (coop_ld_acquire(bar + 1) < gsz) {
                if (++spins > (1u << 27)) __trap();
            }
            bar[0] = 0;
            bar[1] = 0;
            __threadfence();
        }
    }
}

}  // namespace pyitd
