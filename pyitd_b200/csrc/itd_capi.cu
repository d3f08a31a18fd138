// itd_capi.cu -- the extern "C" boundary declared in include/pyitd_b200.h: plan/workspace
// management and the launch sequence of one decomposition.  No torch types, no host syncs on the
// device entry points.
#include "../../include/pyitd_b200.h"

#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>

#include "itd_kernels.cuh"
#include "itd_stream.cuh"
#include "itd_strided.cuh"
#include "itd_sweep.cuh"
#include "itd_coop.cuh"
#include "itd_resident.cuh"
#include "itd_spline.cuh"
#include "itd_sift2d.cuh"
#include "itd_analytics.cuh"

using namespace pyitd;

static thread_local std::string g_err;

static int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}

#define CU(call)                                                                             \
    do {                                                                                     \
        cudaError_t _e = (call);                                                             \
        if (_e != cudaSuccess)                                                               \
            return fail(PYITD_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e));   \
    } while (0)

constexpr int kMaxGroups = 16;

// every entry point runs on the plan's device and puts the caller's current device back on return
struct DeviceGuard {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) err = cudaSetDevice(dev);
    }
    ~DeviceGuard() {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};
#define ON_DEVICE(dev)      \
    DeviceGuard _guard(dev); \
    CU(_guard.err)

struct pyitd_plan {
    int device = 0;
    long long S = 0;          // signals
    int n = 0;                // samples per signal
    int dtype = PYITD_F64;
    int max_iteration = 11, min_extrema = 2;
    unsigned opts = 0;
    int emax = 12, rows = 13;
    int tile_cfg = 1;         // index into the (THREADS, ITEMS) table
    int tile = 1024, tiles = 0;
    bool stream = false;      // one-CTA-per-signal TMA-pipelined level kernel (itd_stream.cuh)
    bool strided = false;     // a few long signals, one at a time: persistent blocks stride over the tiles (itd_strided.cuh)
    // many signals: the whole decomposition in one persistent launch (itd_sweep.cuh).  `stream` stays set when the shape
    // allows it: the single-level entry points (extract_level, find_knots, ...) keep using those kernels.
    bool sweep = false;
    int sw_spans = 0, sw_spw = 0, sw_rs = 0, sw_grid[2] = {0, 0};
    int *sw_ticket = nullptr, *sw_done = nullptr, *sw_rcount[2] = {nullptr, nullptr};
    // a handful of signals: the whole decomposition in one cooperative launch with the signal on chip (itd_coop.cuh)
    bool coop = false;
    int coop_C = 0, coop_gsz = 0, coop_groups = 0;
    int *coop_bar = nullptr;
    void *coop_sum = nullptr;
    // fused pairs of extractions (itd_sweep.cuh): per-CTA scratch for the knots of the baseline between the two
    int sw_mid_ctas = 0;
    unsigned *sw_mid_mask = nullptr;
    unsigned long long *sw_stage_ns = nullptr;
    int strided_cap = 0;      // test hook: upper bound on the persistent grid (PYITD_STRIDED_CTAS)
    // whole-decomposition-on-chip kernel (itd_resident.cuh): one cluster per signal, one launch per batch
    bool resident = false;
    int res_cfg = 0;          // 0: 8 warps x 8 samples/lane, 1: 16 warps x 4 samples/lane
    int res_cl = 1;           // CTAs per cluster
    int res_nu = 0, res_chunk_units = 0, res_clusters = 0;
    size_t res_smem = 0;
    void *res_backup = nullptr;
    int *res_kind = nullptr;

    size_t carry_elem = 8, io_elem = 8;
    // workspace
    void *ws = nullptr;
    size_t ws_bytes = 0;
    void *carry[2] = {nullptr, nullptr};
    KnotTable table[2];
    void *ls = nullptr;       // knot_ls_kernel table [S, lscap + 4] of {L, slope} (stream / strided paths)
    int lscap = 0;
    unsigned long long *desc = nullptr;
    int *stop_e = nullptr, *stop_kind = nullptr, *input_knots = nullptr;
    unsigned tag = 0;
    int launches = 0;
    // stream path: the batch is cut into `groups` signal ranges, each with its own launch chain on its own
    // stream, so that the tail of one range's level e overlaps the head of another range's level e' (a
    // grid of equal-length one-CTA-per-signal blocks otherwise idles most SMs for the last partial wave
    // of every launch)
    int groups = 1;
    cudaStream_t gstream[kMaxGroups] = {};
    cudaEvent_t gfork = nullptr, gjoin[kMaxGroups] = {};
    // optional per-launch CUDA-event timing (bench.py's roofline leg)
    bool timing = false;
    cudaEvent_t *events = nullptr;
    int n_events = 0, events_used = 0;
    // the _host entry point walks the batch in chunks through two slots (sub-plans with their own
    // workspace, device mirrors and stream) so that H2D, the kernels and D2H of neighbouring chunks overlap
    struct HostSlot {
        pyitd_plan *sub = nullptr;
        void *x = nullptr, *rot = nullptr, *bas = nullptr;
        int *ints = nullptr;      // n_rows | knot_counts | input_knots | stop_kind | status
        int *h_rows = nullptr;    // pinned host copy of n_rows (valid-rows-only D2H)
        cudaStream_t stream = nullptr;
    } slot[2];
    long long host_chunk = 0;
    // A plan's workspace (carry buffers, knot tables, look-back descriptors, stop bookkeeping) belongs to ONE call at
    // a time.  `mu` serialises host threads for the duration of a call (calls only enqueue work); `busy` is recorded on
    // the caller's stream at the end of every device call and the next call on a DIFFERENT stream waits for it, so
    // two calls that share a cached plan never overlap on the GPU.
    std::mutex mu;
    cudaEvent_t busy = nullptr;
    cudaStream_t busy_stream = nullptr;
    bool busy_valid = false;
};

// order this call after the plan's previous call when that one ran on another stream.
// A stream that is being CAPTURED into a CUDA graph (the decomposition is a fixed sequence of memsets and one or a few
// launches with no host round trip, so it captures; the workspace must exist: call once before capturing) cannot wait for an
// event recorded outside the capture, nor may an event recorded inside it be waited for outside: the ordering of a graph
// against the plan's other calls is the caller's, as for any captured work.
static bool stream_is_capturing(cudaStream_t st) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return cs != cudaStreamCaptureStatusNone;
}
static int plan_acquire(pyitd_plan *pl, cudaStream_t st) {
    if (pl->busy_valid && pl->busy_stream != st && !stream_is_capturing(st)) CU(cudaStreamWaitEvent(st, pl->busy, 0));
    return 0;
}
static int plan_release(pyitd_plan *pl, cudaStream_t st) {
    if (stream_is_capturing(st)) return 0;
    if (!pl->busy) CU(cudaEventCreateWithFlags(&pl->busy, cudaEventDisableTiming));
    CU(cudaEventRecord(pl->busy, st));
    pl->busy_stream = st;
    pl->busy_valid = true;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// tile configurations.  T = THREADS * ITEMS samples per CTA.
// ---------------------------------------------------------------------------------------------
struct TileCfg {
    int threads, items;
};
static const TileCfg kTileCfgs[] = {{128, 4}, {256, 4}, {256, 8}, {512, 4}};
static const int kNumTileCfgs = 4;

template <typename InT, typename CarryT, int TH, int IT>
static cudaError_t launch_scan_t(const ScanParams &p, long long ctas, cudaStream_t st) {
    auto k = knot_scan_kernel<InT, CarryT, TH, IT>;
    constexpr size_t smem = scan_smem_bytes<TH, IT, CarryT>();
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    k<<<(unsigned)ctas, TH, smem, st>>>(p);
    return cudaGetLastError();
}

template <typename InT, typename CarryT, typename OutT, int TH, int IT>
static cudaError_t launch_level_t(const LevelParams &p, long long ctas, cudaStream_t st) {
    auto k = level_kernel<InT, CarryT, OutT, TH, IT>;
    constexpr size_t smem = level_smem_bytes<TH, IT, CarryT>();
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    k<<<(unsigned)ctas, TH, smem, st>>>(p);
    return cudaGetLastError();
}

template <typename InT, typename CarryT>
static cudaError_t launch_scan_cfg(int cfg, const ScanParams &p, long long ctas, cudaStream_t st) {
    switch (cfg) {
        case 0: return launch_scan_t<InT, CarryT, 128, 4>(p, ctas, st);
        case 1: return launch_scan_t<InT, CarryT, 256, 4>(p, ctas, st);
        case 2: return launch_scan_t<InT, CarryT, 256, 8>(p, ctas, st);
        default: return launch_scan_t<InT, CarryT, 512, 4>(p, ctas, st);
    }
}
template <typename InT, typename CarryT, typename OutT>
static cudaError_t launch_level_cfg(int cfg, const LevelParams &p, long long ctas, cudaStream_t st) {
    switch (cfg) {
        case 0: return launch_level_t<InT, CarryT, OutT, 128, 4>(p, ctas, st);
        case 1: return launch_level_t<InT, CarryT, OutT, 256, 4>(p, ctas, st);
        case 2: return launch_level_t<InT, CarryT, OutT, 256, 8>(p, ctas, st);
        default: return launch_level_t<InT, CarryT, OutT, 512, 4>(p, ctas, st);
    }
}

// one CTA per signal, 8 warps x 4 samples/lane = 1024-sample tiles, 2-stage TMA ring, 3 CTAs per SM
#ifndef PYITD_STREAM_WARPS
#define PYITD_STREAM_WARPS 8          // experiment hook: 4 = 512-sample tiles for the one-CTA-per-signal kernels
#endif
constexpr int kStridedWarps = 8;      // the strided path's passes are written for 1024-sample tiles
constexpr int kStreamWarps = PYITD_STREAM_WARPS, kStreamItems = 4, kStreamStages = 2;
constexpr int kStreamTile = kStreamWarps * 32 * kStreamItems;

template <typename InT, typename CarryT, typename OutT, bool LAST, bool BAS>
static cudaError_t launch_stream_v(const LevelParams &p, long long ctas, cudaStream_t st) {
    auto k = level_stream_kernel<InT, CarryT, OutT, kStreamWarps, kStreamItems, kStreamStages, LAST, BAS>;
    constexpr size_t smem = sizeof(StreamSmem<InT, CarryT, kStreamWarps, kStreamItems, kStreamStages, true>);
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k<<<(unsigned)ctas, kStreamWarps * 32, smem, st>>>(p);
    return cudaGetLastError();
}
template <typename InT, typename CarryT, typename OutT>
static cudaError_t launch_stream_t(const LevelParams &p, long long ctas, cudaStream_t st) {
    const bool last = (p.e == p.emax), bas = (p.bas != nullptr);
    if (last) return bas ? launch_stream_v<InT, CarryT, OutT, true, true>(p, ctas, st)
                         : launch_stream_v<InT, CarryT, OutT, true, false>(p, ctas, st);
    return bas ? launch_stream_v<InT, CarryT, OutT, false, true>(p, ctas, st)
               : launch_stream_v<InT, CarryT, OutT, false, false>(p, ctas, st);
}
template <typename InT, typename CarryT>
static cudaError_t launch_scan_stream_t(const ScanParams &p, long long ctas, cudaStream_t st) {
    auto k = scan_stream_kernel<InT, CarryT, kStreamWarps, kStreamItems, kStreamStages>;
    constexpr size_t smem = sizeof(StreamSmem<InT, CarryT, kStreamWarps, kStreamItems, kStreamStages, false>);
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k<<<(unsigned)ctas, kStreamWarps * 32, smem, st>>>(p);
    return cudaGetLastError();
}



// one long signal: persistent grid of as many blocks as fit on the device (all co-resident: the look-back chain
// over the tiles needs every block to make progress), block c takes tiles c, c + G, ...
template <typename InT, typename CarryT, typename OutT, bool LAST, bool BAS>
static cudaError_t launch_strided_v(const LevelParams &p, int cap, cudaStream_t st) {
    auto k = level_strided_kernel<InT, CarryT, OutT, kStridedWarps, kStreamItems, kStreamStages, LAST, BAS>;
    constexpr size_t smem = sizeof(StridedSmem<InT, CarryT, kStridedWarps, kStreamItems, kStreamStages>);
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0, dev = 0, sms = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, kStridedWarps * 32, smem);
    if (e != cudaSuccess) return e;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long g = (long long)per_sm * sms;
    if (g < 1) return cudaErrorLaunchOutOfResources;
    if (cap > 0 && g > cap) g = cap;
    if (g > p.tiles) g = p.tiles;
    k<<<(unsigned)g, kStridedWarps * 32, smem, st>>>(p);
    return cudaGetLastError();
}
template <typename InT, typename CarryT, typename OutT>
static cudaError_t launch_strided_t(const LevelParams &p, int cap, cudaStream_t st) {
    const bool last = (p.e == p.emax), bas = (p.bas != nullptr);
    if (last) return bas ? launch_strided_v<InT, CarryT, OutT, true, true>(p, cap, st)
                         : launch_strided_v<InT, CarryT, OutT, true, false>(p, cap, st);
    return bas ? launch_strided_v<InT, CarryT, OutT, false, true>(p, cap, st)
               : launch_strided_v<InT, CarryT, OutT, false, false>(p, cap, st);
}

// place_knots_kernel: a warp owns a group of 32 tiles, 8 warps per block; 4 blocks per SM measured best on 2^28 samples
// (profiles/r1/s5/place_sweep.log: more blocks only widen the window of DRAM pages in flight)
static int compact_grid(int tiles) {
    const int groups = (tiles + 31) / 32, blocks = (groups + 7) / 8;
    static const int cap = getenv("PYITD_PLACE_BLOCKS") ? atoi(getenv("PYITD_PLACE_BLOCKS")) : 148 * 4;   // test hook
    return blocks < cap ? blocks : cap;
}

// the strided path's knot scan: flag words + per-tile counts, then the prefix and the compaction pass
template <typename InT, typename CarryT>
static cudaError_t launch_scan_strided_t(const pyitd_plan *pl, const ScanParams &p, cudaStream_t st) {
    auto k = scan_strided_kernel<InT, CarryT, kStridedWarps, kStreamItems>;
    constexpr size_t smem = sizeof(ScanStridedSmem<InT, kStridedWarps, kStreamItems>);
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0, dev = 0, sms = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, kStridedWarps * 32, smem);
    if (e != cudaSuccess) return e;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long g = (long long)per_sm * sms;
    if (g < 1) return cudaErrorLaunchOutOfResources;
    if (pl->strided_cap > 0 && g > pl->strided_cap) g = pl->strided_cap;
    if (g > p.tiles) g = p.tiles;
    k<<<(unsigned)g, kStridedWarps * 32, smem, st>>>(p);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    tile_prefix_scan_kernel<InT, CarryT><<<1, 1024, 0, st>>>(p.out, p.x, p.sig0, p.tiles, p.n, p.input_knots);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    place_knots_kernel<CarryT><<<compact_grid(p.tiles), 256, 0, st>>>(p.out, p.sig0, p.tiles, -1, nullptr);
    return cudaGetLastError();
}

static bool strided_launchable_fwd(const pyitd_plan *pl, const void *in) {
    return pl->strided && (reinterpret_cast<uintptr_t>(in) & 15) == 0;
}

static bool stream_launchable(const pyitd_plan *pl, const void *in) {
    return pl->stream && (reinterpret_cast<uintptr_t>(in) & 15) == 0;
}

// nsig: signals of this launch (stream path only: p.sig0 .. p.sig0 + nsig); the look-back path always
// launches the whole batch
static cudaError_t launch_scan(const pyitd_plan *pl, const ScanParams &p, cudaStream_t st, long long nsig) {
    if (strided_launchable_fwd(pl, p.x)) {
        // a few long signals: one after the other, each launch fills the device
        ScanParams q = p;
        // the per-group knot sums start from zero (a single-level call leaves its level kernel's sums unconsumed)
        cudaError_t ze = cudaMemsetAsync(p.out.gsum, 0, (size_t)pl->S * (size_t)p.out.gstride * sizeof(int), st);
        if (ze != cudaSuccess) return ze;
        for (long long sg = 0; sg < pl->S; ++sg) {
            q.sig0 = (int)sg;
            cudaError_t e;
            switch (pl->dtype) {
                case PYITD_F64: e = launch_scan_strided_t<double, double>(pl, q, st); break;
                case PYITD_F32_MIXED: e = launch_scan_strided_t<float, double>(pl, q, st); break;
                default: e = launch_scan_strided_t<float, float>(pl, q, st);
            }
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
    if (stream_launchable(pl, p.x)) {
        switch (pl->dtype) {
            case PYITD_F64: return launch_scan_stream_t<double, double>(p, nsig, st);
            case PYITD_F32_MIXED: return launch_scan_stream_t<float, double>(p, nsig, st);
            default: return launch_scan_stream_t<float, float>(p, nsig, st);
        }
    }
    const long long ctas = pl->S * pl->tiles;
    switch (pl->dtype) {
        case PYITD_F64: return launch_scan_cfg<double, double>(pl->tile_cfg, p, ctas, st);
        case PYITD_F32_MIXED: return launch_scan_cfg<float, double>(pl->tile_cfg, p, ctas, st);
        default: return launch_scan_cfg<float, float>(pl->tile_cfg, p, ctas, st);
    }
}
// first = the launch reads the caller's input (io type) instead of a carry buffer
static cudaError_t launch_level(const pyitd_plan *pl, const LevelParams &p, bool first, cudaStream_t st,
                                long long nsig) {
    if (strided_launchable_fwd(pl, p.in)) {
        LevelParams q = p;
        // a block takes a contiguous run of tiles (measured 19.5 -> 18.4 ms on 2^28 samples: the knot-free-span cache then
        // survives from tile to tile on deep levels); PYITD_STRIDED_CONTIG=0 restores every-G-th-tile striding
        const bool contig = !(getenv("PYITD_STRIDED_CONTIG") && atoi(getenv("PYITD_STRIDED_CONTIG")) == 0);
        if (contig) q.opts |= kOptContigTiles;
        for (long long sg = 0; sg < pl->S; ++sg) {
            q.sig0 = (int)sg;
            cudaError_t e;
            switch (pl->dtype) {
                case PYITD_F64: e = launch_strided_t<double, double, double>(q, pl->strided_cap, st); break;
                case PYITD_F32_MIXED:
                    e = first ? launch_strided_t<float, double, float>(q, pl->strided_cap, st)
                              : launch_strided_t<double, double, float>(q, pl->strided_cap, st);
                    break;
                default: e = launch_strided_t<float, float, float>(q, pl->strided_cap, st);
            }
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
    if (stream_launchable(pl, p.in)) {
        switch (pl->dtype) {
            case PYITD_F64: return launch_stream_t<double, double, double>(p, nsig, st);
            case PYITD_F32_MIXED:
                return first ? launch_stream_t<float, double, float>(p, nsig, st)
                             : launch_stream_t<double, double, float>(p, nsig, st);
            default: return launch_stream_t<float, float, float>(p, nsig, st);
        }
    }
    const long long ctas = pl->S * pl->tiles;
    switch (pl->dtype) {
        case PYITD_F64: return launch_level_cfg<double, double, double>(pl->tile_cfg, p, ctas, st);
        case PYITD_F32_MIXED:
            return first ? launch_level_cfg<float, double, float>(pl->tile_cfg, p, ctas, st)
                         : launch_level_cfg<double, double, float>(pl->tile_cfg, p, ctas, st);
        default: return launch_level_cfg<float, float, float>(pl->tile_cfg, p, ctas, st);
    }
}

// ---------------------------------------------------------------------------------------------
// resident kernel: configuration and launch
// ---------------------------------------------------------------------------------------------
constexpr size_t kMaxSmemOptin = 227 * 1024;

template <typename InT, typename CarryT, typename OutT, int W, int SPL>
static cudaError_t res_launch_t(const ResidentParams &p, int cl, int clusters, size_t smem, cudaStream_t st,
                                int *max_clusters_out) {
    auto k = resident_kernel<InT, CarryT, OutT, W, SPL>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(W * 32, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (max_clusters_out) {
        cfg.gridDim = dim3((unsigned)cl, 1, 1);
        int nc = 0;
        e = cudaOccupancyMaxActiveClusters(&nc, k, &cfg);
        if (e != cudaSuccess) return e;
        *max_clusters_out = nc;
        return cudaSuccess;
    }
    cfg.gridDim = dim3((unsigned)(clusters * cl), 1, 1);
    return cudaLaunchKernelEx(&cfg, k, p);
}

template <int W, int SPL>
static cudaError_t res_launch_d(int dtype, const ResidentParams &p, int cl, int clusters, size_t smem,
                                cudaStream_t st, int *mc) {
    switch (dtype) {
        case PYITD_F64: return res_launch_t<double, double, double, W, SPL>(p, cl, clusters, smem, st, mc);
        case PYITD_F32_MIXED: return res_launch_t<float, double, float, W, SPL>(p, cl, clusters, smem, st, mc);
        default: return res_launch_t<float, float, float, W, SPL>(p, cl, clusters, smem, st, mc);
    }
}
static cudaError_t res_launch(const pyitd_plan *pl, const ResidentParams &p, int clusters, cudaStream_t st, int *mc) {
    if (pl->res_cfg == 0) return res_launch_d<8, 8>(pl->dtype, p, pl->res_cl, clusters, pl->res_smem, st, mc);
    return res_launch_d<16, 4>(pl->dtype, p, pl->res_cl, clusters, pl->res_smem, st, mc);
}

// shared-memory layout of configuration (cfg, cl) for n samples; returns false when it does not fit
static bool res_geometry(int cfg, int dtype, int n, int cl, ResidentParams &p, size_t &smem) {
    const int W = cfg == 0 ? 8 : 16, SPL = cfg == 0 ? 8 : 4, UNIT = 32 * SPL;
    if (cl - 1 + W > 32) return false;
    const int nu = (n + UNIT - 1) / UNIT, gw = cl * W;
    int chunk = 1;
    for (int r = 0; r < cl; ++r) {
        const int c = res_unit_begin((r + 1) * W, gw, nu) - res_unit_begin(r * W, gw, nu);
        if (c > chunk) chunk = c;
    }
    p.nu = nu;
    p.chunk_units = chunk;
    if (dtype == PYITD_F32) {
        if (cfg == 0) ResidentGeom<float, 8, 8>::layout(chunk, p, smem);
        else ResidentGeom<float, 16, 4>::layout(chunk, p, smem);
    } else {
        if (cfg == 0) ResidentGeom<double, 8, 8>::layout(chunk, p, smem);
        else ResidentGeom<double, 16, 4>::layout(chunk, p, smem);
    }
    return smem <= kMaxSmemOptin;
}

// picks the smallest cluster that holds the signal on chip; false when none does
static bool res_configure(pyitd_plan *pl) {
    int cfg = 1;                               // 16 warps x 4 samples per lane measured faster than 8 x 8
    if (const char *env = getenv("PYITD_RES_CFG")) cfg = atoi(env) ? 1 : 0;
    int force_cl = 0;
    if (const char *env = getenv("PYITD_RES_CL")) force_cl = atoi(env);
    for (int cl = 1; cl <= 8; cl *= 2) {
        if (force_cl && cl != force_cl) continue;
        ResidentParams p = {};
        size_t smem = 0;
        if (!res_geometry(cfg, pl->dtype, pl->n, cl, p, smem)) continue;
        // a cluster wider than the signal has units would leave CTAs without work: still correct, but pointless
        pl->res_cfg = cfg;
        pl->res_cl = cl;
        pl->res_nu = p.nu;
        pl->res_chunk_units = p.chunk_units;
        pl->res_smem = smem;
        return true;
    }
    return false;
}

// ---------------------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------------------
static size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

extern "C" int pyitd_abi_version(void) { return PYITD_ABI_VERSION; }
extern "C" const char *pyitd_last_error(void) { return g_err.c_str(); }
extern "C" int pyitd_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int pyitd_plan_create(pyitd_plan **out, int device, int64_t n_signals, int64_t n_samples,
                                 int dtype, int max_iteration, int min_extrema, int options) {
    if (!out) return fail(PYITD_E_INVALID, "plan pointer is null");
    *out = nullptr;
    if (n_signals < 1) return fail(PYITD_E_INVALID, "n_signals must be >= 1");
    if (n_samples < 3) return fail(PYITD_E_INVALID, "n_samples must be >= 3 (ITD.py:42-43 is undefined below 3)");
    if (n_samples > 0x7ffffff0ll) return fail(PYITD_E_INVALID, "n_samples must be < 2^31");
    if (dtype != PYITD_F64 && dtype != PYITD_F32_MIXED && dtype != PYITD_F32)
        return fail(PYITD_E_INVALID, "unknown dtype");
    if (max_iteration < 0 || max_iteration > 4096) return fail(PYITD_E_INVALID, "max_iteration out of range");
    if (min_extrema < 0) return fail(PYITD_E_INVALID, "min_extrema must be >= 0");
    int ndev = pyitd_device_count();
    if (ndev <= 0) return fail(PYITD_E_NODEVICE, "no CUDA device visible");
    if (device < 0 || device >= ndev) return fail(PYITD_E_INVALID, "device index out of range");
    ON_DEVICE(device);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(PYITD_E_NODEVICE, std::string("device ") + prop.name + " is not sm_100-class; this library ships sm_100a code only");

    pyitd_plan *pl = new (std::nothrow) pyitd_plan();
    if (!pl) return fail(PYITD_E_NOMEM, "host allocation failed");
    pl->device = device;
    pl->S = n_signals;
    pl->n = (int)n_samples;
    pl->dtype = dtype;
    pl->max_iteration = max_iteration;
    pl->min_extrema = min_extrema;
    pl->opts = (unsigned)options;
    pl->emax = max_iteration + 1;
    pl->rows = max_iteration + 2;
    pl->carry_elem = (dtype == PYITD_F32) ? 4 : 8;
    pl->io_elem = (dtype == PYITD_F64) ? 8 : 4;

    // look-back kernels: 512 x 4 tiles measured 15 % faster than 256 x 4 on long signals (profiles/r1/s3/cfg3_tiles.jsonl)
    int cfg = (n_samples <= 512) ? 0 : ((n_samples >= (1 << 20)) ? 3 : 1);
    if (const char *env = getenv("PYITD_TILE_CFG")) {
        int v = atoi(env);
        if (v >= 0 && v < kNumTileCfgs) cfg = v;
    }
    // path choice (measured on B200, profiles/r1/path_sweep_r1.json):
    //   many signals            -> one TMA-pipelined CTA per signal, carry in HBM (stream)
    //   a few dozen signals     -> one cluster per signal, carry on chip, one launch (resident)
    //   a handful / long / tiny -> many look-back CTAs per signal (lookback)
    // The streaming kernel needs 16-byte aligned rows for its TMA bulk copies.
    const long long stream_tiles = (n_samples + kStreamTile - 1) / kStreamTile;
    const bool stream_ok_rows = (n_samples % 4 == 0);             // 16-byte aligned rows for the TMA bulk copies
    const bool stream_ok = stream_ok_rows && stream_tiles <= kStreamMaxTiles;
    bool stream = stream_ok && n_samples >= 2048 && n_signals >= 160;
    // geometry of the cooperative kernel (itd_coop.cuh): chunk = about one CTA per SM for one signal; how many signals the
    // chip holds at once (its shared memory takes ~9 signals of 65 536 fp64 samples, ~37 of 8192)
    int coop_sms = 0;
    cudaDeviceGetAttribute(&coop_sms, cudaDevAttrMultiProcessorCount, device);
    long long coop_C = (n_samples + coop_sms - 1) / coop_sms;
    coop_C = ((coop_C + 255) / 256) * 256;
    if (const char *env = getenv("PYITD_COOP_CHUNK")) {
        const long long v = atoll(env);
        if (v >= 256 && v % 256 == 0) coop_C = v;
    }
    const long long coop_gsz = (n_samples + coop_C - 1) / coop_C;
    bool coop_one_round = false;
    if (coop_C <= kCoopMaxChunk && n_samples >= 3) {
        const size_t csm = (pl->carry_elem == 8) ? coop_smem_bytes<double>((int)coop_C) : coop_smem_bytes<float>((int)coop_C);
        long long per_sm = (long long)(227 * 1024) / (long long)(csm + 1024);
        if (per_sm > 2048 / kCoopThreads) per_sm = 2048 / kCoopThreads;
        // measured (profiles/r2/mid_batch_probe.jsonl, coop_shapes.jsonl): one round of the cooperative kernel beats the
        // cluster-resident kernel up to ~24 signals (8192 samples: 61 vs 106 us at 17 signals, 104 vs 107 at 32); a second
        // round does not (65 536 samples: 292 us for 16 signals in two rounds vs 209 us resident)
        coop_one_round = n_signals <= per_sm * coop_sms / coop_gsz && n_signals <= 24;
    }
    bool resident = !stream && !coop_one_round && n_signals > 8 && n_signals < 160 && n_samples >= 4096;
    if (const char *env = getenv("PYITD_FORCE_PATH")) {
        stream = !strcmp(env, "stream") && stream_ok;
        resident = !strcmp(env, "resident");
    }
    pl->resident = resident && res_configure(pl);
    // one long signal: the streaming pipeline made persistent over its tiles (2.4x the look-back kernel at 2^28)
    bool strided = !stream && !pl->resident && n_signals <= 16 && stream_ok_rows && n_samples >= (1 << 18);
    if (const char *env = getenv("PYITD_FORCE_PATH")) strided = !strcmp(env, "strided") && n_signals <= 64 && stream_ok_rows;
    if (pl->resident) stream = strided = false;
    if (stream || strided) cfg = (stream && kStreamWarps == 4) ? 0 : 1;   // the look-back kernels must agree on the tile of the stream / strided kernels
    pl->strided = strided;
    if (const char *env = getenv("PYITD_STRIDED_CTAS")) pl->strided_cap = atoi(env);
    pl->stream = stream;
    // the sweep kernel takes over pyitd_decompose_* wherever the one-CTA-per-signal kernels were chosen (it has no row
    // alignment requirement: plain 8-byte loads); PYITD_FORCE_PATH=stream keeps the round-1 launch chain
    bool sweep = !pl->resident && !strided && n_signals >= 160 && n_samples >= 2048;
    if (const char *env = getenv("PYITD_FORCE_PATH")) sweep = !strcmp(env, "sweep");
    pl->sweep = sweep;
    // up to 16 signals that the look-back kernels would take: one cooperative launch, signal on chip, one group barrier per
    // extraction (PYITD_FORCE_PATH=lookback keeps the launch chain).  Chunk: about one CTA per SM for one signal.
    {
        bool coop = !stream && !pl->resident && !strided && !sweep && (coop_one_round || n_signals <= 16) && n_samples >= 3 &&
                    coop_C <= kCoopMaxChunk;
        if (const char *env = getenv("PYITD_FORCE_PATH"))
            coop = !strcmp(env, "coop") && n_signals <= 64 && n_samples >= 3 && coop_C <= kCoopMaxChunk && !pl->resident;
        pl->coop = coop;
        pl->coop_C = (int)coop_C;
        pl->coop_gsz = (int)coop_gsz;
    }
    pl->sw_spans = (int)((n_samples + kSweepSpan - 1) / kSweepSpan);
    pl->sw_spw = (pl->sw_spans + kSweepWarps - 1) / kSweepWarps;
    pl->sw_rs = pl->sw_spw * kSweepSpan + 8;
    // two launch chains hide most of the partial last wave of every level launch (measured: 17.35 -> 16.56 ms/step
    // on 4096 x 65536; more groups add nothing)
    pl->groups = (stream && !sweep && n_signals >= 1024) ? 2 : 1;
    if (const char *env = getenv("PYITD_GROUPS")) {
        const int v = atoi(env);
        if (v >= 1 && v <= kMaxGroups) pl->groups = v;
    }
    pl->tile_cfg = cfg;
    pl->tile = kTileCfgs[cfg].threads * kTileCfgs[cfg].items;
    pl->tiles = (int)((n_samples + pl->tile - 1) / pl->tile);
    if ((long long)pl->tiles * pl->S > 0x7fffffffll) {
        delete pl;
        return fail(PYITD_E_INVALID, "n_signals * tiles exceeds the grid limit; split the batch");
    }

    *out = pl;
    return 0;
}

// the device workspace is allocated on the first device call: a plan that is only used through
// pyitd_decompose_host never needs the full-batch workspace (its chunk sub-plans own theirs)
static int ensure_workspace(pyitd_plan *pl, cudaStream_t st) {
    if (pl->ws) return 0;
    const size_t SN = (size_t)pl->S * (size_t)pl->n;
    long long kstride = (((long long)pl->n + 3) & ~3ll) + 4;
    if (pl->sweep && kstride < (long long)kSweepWarps * pl->sw_rs) kstride = (long long)kSweepWarps * pl->sw_rs;   // eight region lists
    const long long mstride = ((((long long)pl->n + 31) / 32) + 3) & ~3ll;
    const size_t SK = (size_t)pl->S * (size_t)kstride;
    const size_t b_carry = align_up(SN * pl->carry_elem);
    const size_t b_xk = align_up(SK * pl->carry_elem);
    const size_t b_tau = align_up(SK * sizeof(int));
    const size_t b_mask = align_up((size_t)pl->S * (size_t)mstride * sizeof(unsigned));
    const size_t b_tbase = align_up((size_t)pl->S * (pl->tiles + 1) * sizeof(int));
    const size_t b_sig = align_up((size_t)pl->S * sizeof(int));
    const size_t b_endl = align_up((size_t)pl->S * 2 * pl->carry_elem);
    const size_t b_desc = align_up((size_t)pl->S * pl->tiles * sizeof(unsigned long long));
    const long long gstride = pl->strided ? ((((long long)pl->tiles + 31) / 32 + 1 + 3) & ~3ll) : 0;
    const size_t b_group = align_up((size_t)pl->S * (size_t)gstride * sizeof(int));
    const size_t b_stage = pl->strided ? b_tau + b_xk : 0;             // tile-local knot staging of the strided path
    // knot baseline table of knot_ls_kernel: signals with at most n / 16 interior knots (sparser tables are where every
    // warp would recompute the same few knots); PYITD_LS=0 turns the pre-pass off
    const bool ls_on = !(getenv("PYITD_LS") && atoi(getenv("PYITD_LS")) == 0);      // read per plan (tests toggle it)
    const int ls_div = getenv("PYITD_LS_DIV") ? (atoi(getenv("PYITD_LS_DIV")) > 1 ? atoi(getenv("PYITD_LS_DIV")) : 16) : 16;   // experiment hook (n/16 measured best: profiles/r1/s5/ls_probe2.log)
    pl->lscap = (ls_on && (pl->stream || pl->strided)) ? (int)((pl->n / ls_div) & ~1ll) : 0;   // even: float rows stay 16-byte aligned
    const size_t b_ls = pl->lscap ? align_up((size_t)pl->S * (size_t)(pl->lscap + 4) * 2 * pl->carry_elem) : 0;
    // sweep path: ticket slots, done[], region counts, stage clocks, and the per-CTA flag-word scratch of the fused pairs
    // (PYITD_SWEEP_FUSE=0 turns them off): 4 CTAs per SM is the kernel's launch bound
    int mid_ctas = 0;
    if (pl->sweep && !(getenv("PYITD_SWEEP_FUSE") && atoi(getenv("PYITD_SWEEP_FUSE")) == 0)) {
        int sms = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, pl->device);
        mid_ctas = 4 * sms;
        if ((long long)mid_ctas > pl->S) mid_ctas = (int)pl->S;
    }
    const size_t b_mid = align_up((size_t)mid_ctas * (size_t)mstride * sizeof(unsigned));
    const size_t b_sweep = pl->sweep ? align_up((size_t)(pl->rows + 4) * sizeof(int)) + align_up((size_t)pl->S * 2 * sizeof(int)) +
                                           2 * align_up((size_t)pl->S * kSweepWarps * sizeof(int)) +
                                           align_up((size_t)(pl->rows + 2) * sizeof(unsigned long long)) + b_mid
                                     : 0;
    size_t total = 2 * b_carry + 2 * (b_tau + b_xk + b_tbase + b_sig + b_endl + b_mask) + b_desc + 3 * b_sig + 2 * b_group + b_stage + b_ls + b_sweep;
    pl->ws_bytes = total;
    cudaError_t ce = cudaMalloc(&pl->ws, total);
    if (ce != cudaSuccess) {
        cudaGetLastError();
        pl->ws = nullptr;
        return fail(PYITD_E_NOMEM, "cudaMalloc of " + std::to_string(total) + " workspace bytes failed: " + cudaGetErrorString(ce));
    }
    char *c = (char *)pl->ws;
    auto take = [&](size_t b) { char *r = c; c += b; return (void *)r; };
    pl->carry[0] = take(b_carry);
    pl->carry[1] = take(b_carry);
    for (int i = 0; i < 2; ++i) {
        pl->table[i].tau = (int *)take(b_tau);
        pl->table[i].xk = take(b_xk);
        pl->table[i].mask = (unsigned *)take(b_mask);
        pl->table[i].kstride = kstride;
        pl->table[i].mstride = mstride;
        pl->table[i].tbase = (int *)take(b_tbase);
        pl->table[i].kcount = (int *)take(b_sig);
        pl->table[i].endl = take(b_endl);
    }
    pl->desc = (unsigned long long *)take(b_desc);
    pl->stop_e = (int *)take(b_sig);
    pl->stop_kind = (int *)take(b_sig);
    pl->input_knots = (int *)take(b_sig);
    {
        int *gsum = (int *)take(b_group), *gbase = (int *)take(b_group);
        int *stau = pl->strided ? (int *)take(b_tau) : nullptr;
        void *sxk = pl->strided ? take(b_xk) : nullptr;
        for (int i = 0; i < 2; ++i) {
            pl->table[i].stau = stau;
            pl->table[i].sxk = sxk;
            pl->table[i].gsum = pl->strided ? gsum : nullptr;
            pl->table[i].gbase = pl->strided ? gbase : nullptr;
            pl->table[i].gstride = gstride;
        }
    }
    pl->ls = b_ls ? take(b_ls) : nullptr;
    if (pl->sweep) {
        pl->sw_ticket = (int *)take(align_up((size_t)(pl->rows + 4) * sizeof(int)));
        pl->sw_done = (int *)take(align_up((size_t)pl->S * 2 * sizeof(int)));              // done[S], sel[S]
        for (int i = 0; i < 2; ++i) pl->sw_rcount[i] = (int *)take(align_up((size_t)pl->S * kSweepWarps * sizeof(int)));
        pl->sw_stage_ns = (unsigned long long *)take(align_up((size_t)(pl->rows + 2) * sizeof(unsigned long long)));
        pl->sw_mid_ctas = mid_ctas;
        if (mid_ctas) pl->sw_mid_mask = (unsigned *)take(b_mid);
    }
    // the mask rows are padded to 4 words: the padding (and everything else) starts out as "no knot"
    // on the CALLER's stream: a cudaStreamNonBlocking stream is not ordered after the legacy default stream, so a
    // synchronous cudaMemset could land after the first kernels of this call had written mask words or descriptors
    ce = cudaMemsetAsync(pl->table[0].mask, 0, b_mask, st);
    if (ce == cudaSuccess) ce = cudaMemsetAsync(pl->table[1].mask, 0, b_mask, st);
    if (ce == cudaSuccess) ce = cudaMemsetAsync(pl->desc, 0, b_desc, st);
    if (ce != cudaSuccess) {
        cudaFree(pl->ws);
        pl->ws = nullptr;
        return fail(PYITD_E_CUDA, std::string("cudaMemset: ") + cudaGetErrorString(ce));
    }
    return 0;
}

extern "C" void pyitd_plan_destroy(pyitd_plan *pl) {
    if (!pl) return;
    DeviceGuard guard(pl->device);
    if (pl->busy) cudaEventDestroy(pl->busy);
    for (auto &sl : pl->slot) {
        if (sl.sub) pyitd_plan_destroy(sl.sub);
        cudaFree(sl.x);
        cudaFree(sl.rot);
        cudaFree(sl.bas);
        cudaFree(sl.ints);
        if (sl.h_rows) cudaFreeHost(sl.h_rows);
        if (sl.stream) cudaStreamDestroy(sl.stream);
    }
    if (pl->events) {
        for (int i = 0; i < pl->n_events; ++i) cudaEventDestroy(pl->events[i]);
        delete[] pl->events;
    }
    for (int i = 0; i < kMaxGroups; ++i) {
        if (pl->gstream[i]) cudaStreamDestroy(pl->gstream[i]);
        if (pl->gjoin[i]) cudaEventDestroy(pl->gjoin[i]);
    }
    if (pl->gfork) cudaEventDestroy(pl->gfork);
    cudaFree(pl->res_backup);
    cudaFree(pl->res_kind);
    cudaFree(pl->coop_bar);
    cudaFree(pl->ws);
    delete pl;
}

extern "C" int pyitd_plan_rows(const pyitd_plan *pl) { return pl ? pl->rows : PYITD_E_INVALID; }
extern "C" int64_t pyitd_plan_workspace_bytes(const pyitd_plan *pl) { return pl ? (int64_t)pl->ws_bytes : 0; }
extern "C" int pyitd_plan_launches(const pyitd_plan *pl) { return pl ? pl->launches : 0; }
extern "C" int pyitd_plan_path(const pyitd_plan *pl, int *cluster_size) {
    if (!pl) return PYITD_E_INVALID;
    if (cluster_size) *cluster_size = pl->resident ? pl->res_cl : 1;
    if (pl->strided) return PYITD_PATH_STRIDED;
    if (pl->sweep) return PYITD_PATH_SWEEP;
    if (pl->coop) return PYITD_PATH_COOP;
    return pl->resident ? PYITD_PATH_RESIDENT : (pl->stream ? PYITD_PATH_STREAM : PYITD_PATH_LOOKBACK);
}

// a fresh look-back tag for every launch; descriptors are only cleared when the 30-bit tag wraps
static int next_tag(pyitd_plan *pl, cudaStream_t st, unsigned *tag) {
    if (pl->tag >= (1u << 30) - 2u) {
        CU(cudaMemsetAsync(pl->desc, 0, (size_t)pl->S * pl->tiles * sizeof(unsigned long long), st));
        pl->tag = 0;
    }
    *tag = ++pl->tag;
    return 0;
}

// event i is recorded before launch i, event i+1 after it, on the launching stream
static int mark(pyitd_plan *pl, cudaStream_t st) {
    if (!pl->timing) return 0;
    if (pl->events_used < pl->n_events) CU(cudaEventRecord(pl->events[pl->events_used++], st));
    return 0;
}

static int run_scan(pyitd_plan *pl, const void *x, int *status, int *input_knots, cudaStream_t st, int kinds = 3,
                    long long sig0 = 0, long long nsig = -1, bool timed = true) {
    ScanParams sp;
    sp.x = x;
    sp.out = pl->table[0];
    sp.desc = pl->desc;
    if (int rc = next_tag(pl, st, &sp.tag)) return rc;
    sp.status = status;
    sp.input_knots = input_knots;
    sp.n = pl->n;
    sp.tiles = pl->tiles;
    sp.kinds = kinds;
    sp.sig0 = (int)sig0;
    if (timed)
        if (int rc = mark(pl, st)) return rc;
    CU(launch_scan(pl, sp, st, nsig < 0 ? pl->S : nsig));
    pl->launches++;
    return timed ? mark(pl, st) : 0;
}

static bool strided_launchable(const pyitd_plan *pl, const void *in) {
    return pl->strided && (reinterpret_cast<uintptr_t>(in) & 15) == 0;
}
// tile_prefix_kernel + place_knots_kernel on the table the level launch `lp` just produced
static int run_strided_passes(pyitd_plan *pl, const LevelParams &lp, cudaStream_t st) {
    const int last = (lp.e == lp.emax) ? 1 : 0;
    const int grid = compact_grid(pl->tiles);
    for (long long sg = 0; sg < pl->S; ++sg) {
    const int sig0 = (int)sg;
    if (pl->carry_elem == 8) {
        tile_prefix_kernel<double><<<1, 1024, 0, st>>>(lp.next, sig0, lp.tiles, lp.n, lp.e, lp.rows, lp.min_extrema, last,
                                                       lp.stop_e, lp.stop_kind, lp.n_rows, lp.knot_counts);
        CU(cudaGetLastError());
        place_knots_kernel<double><<<grid, 256, 0, st>>>(lp.next, sig0, lp.tiles, lp.e, lp.stop_e);
    } else {
        tile_prefix_kernel<float><<<1, 1024, 0, st>>>(lp.next, sig0, lp.tiles, lp.n, lp.e, lp.rows, lp.min_extrema, last,
                                                      lp.stop_e, lp.stop_kind, lp.n_rows, lp.knot_counts);
        CU(cudaGetLastError());
        place_knots_kernel<float><<<grid, 256, 0, st>>>(lp.next, sig0, lp.tiles, lp.e, lp.stop_e);
    }
    CU(cudaGetLastError());
    pl->launches += 2;
    }
    return 0;
}

// signal ranges of the stream path's launch groups
static int effective_groups(const pyitd_plan *pl, const void *x) {
    if (!stream_launchable(pl, x)) return 1;
    int g = pl->groups;
    if (g > kMaxGroups) g = kMaxGroups;
    if ((long long)g > pl->S) g = (int)pl->S;
    return g < 1 ? 1 : g;
}
static int ensure_group_streams(pyitd_plan *pl, int g) {
    if (!pl->gfork) CU(cudaEventCreateWithFlags(&pl->gfork, cudaEventDisableTiming));
    for (int i = 0; i < g; ++i) {
        if (!pl->gstream[i]) CU(cudaStreamCreateWithFlags(&pl->gstream[i], cudaStreamNonBlocking));
        if (!pl->gjoin[i]) CU(cudaEventCreateWithFlags(&pl->gjoin[i], cudaEventDisableTiming));
    }
    return 0;
}

static int run_resident(pyitd_plan *pl, const void *x, void *rotations, void *baselines, int32_t *n_rows,
                        int32_t *knot_counts, int32_t *input_knots, int32_t *stop_kind, int32_t *status,
                        cudaStream_t st) {
    ResidentParams rp = {};
    size_t smem = 0;
    if (!res_geometry(pl->res_cfg, pl->dtype, pl->n, pl->res_cl, rp, smem))
        return fail(PYITD_E_INVALID, "resident configuration no longer fits");
    const int unit = pl->res_cfg == 0 ? 256 : 128;
    if (!pl->res_clusters) {
        int mc = 0;
        CU(res_launch(pl, rp, 0, st, &mc));
        if (mc < 1) return fail(PYITD_E_CUDA, "no resident cluster fits on this device");
        if (const char *env = getenv("PYITD_RES_CLUSTERS")) {
            const int v = atoi(env);
            if (v >= 1 && v < mc) mc = v;
        }
        pl->res_clusters = (int)((long long)mc < pl->S ? mc : pl->S);
        const size_t b_backup = (size_t)pl->res_clusters * rp.nu * unit * pl->carry_elem;
        cudaError_t ce = cudaMalloc(&pl->res_backup, b_backup);
        if (ce == cudaSuccess) ce = cudaMalloc((void **)&pl->res_kind, (size_t)pl->S * sizeof(int));
        if (ce != cudaSuccess) {
            cudaGetLastError();
            pl->res_clusters = 0;
            return fail(PYITD_E_NOMEM, std::string("resident scratch allocation failed: ") + cudaGetErrorString(ce));
        }
        pl->ws_bytes = b_backup + (size_t)pl->S * sizeof(int);
    }
    pl->launches = 0;
    pl->events_used = 0;
    const size_t b_sig = (size_t)pl->S * sizeof(int);
    CU(cudaMemsetAsync(status, 0, b_sig, st));
    CU(cudaMemsetAsync(knot_counts, 0, b_sig * pl->rows, st));
    rp.x = x;
    rp.rot = rotations;
    rp.bas = baselines;
    rp.backup = pl->res_backup;
    rp.out_sig_stride = (long long)pl->rows * pl->n;
    rp.backup_stride = (long long)rp.nu * unit;
    rp.n_rows = n_rows;
    rp.knot_counts = knot_counts;
    rp.input_knots = input_knots;
    rp.stop_kind = stop_kind ? stop_kind : pl->res_kind;
    rp.status = status;
    rp.S = pl->S;
    rp.n = pl->n;
    rp.emax = pl->emax;
    rp.rows = pl->rows;
    rp.min_extrema = pl->min_extrema;
    rp.opts = pl->opts;
    if (int rc = mark(pl, st)) return rc;
    CU(res_launch(pl, rp, pl->res_clusters, st, nullptr));
    pl->launches = 1;
    return mark(pl, st);
}

static int run_sweep(pyitd_plan *pl, const void *x, void *rotations, void *baselines, int32_t *n_rows,
                     int32_t *knot_counts, int32_t *input_knots, int *sk, int32_t *status, cudaStream_t st);

// ---------------------------------------------------------------------------------------------
// cooperative kernel: a handful of signals, each kept on chip by a group of CTAs (itd_coop.cuh)
// ---------------------------------------------------------------------------------------------
constexpr int kCoopNoFit = 12345;
template <typename InT, typename CarryT, typename OutT>
static int coop_launch_t(pyitd_plan *pl, CoopParams &p, bool bas, cudaStream_t st) {
    auto k = bas ? coop_kernel<InT, CarryT, OutT, true> : coop_kernel<InT, CarryT, OutT, false>;
    const size_t smem = coop_smem_bytes<CarryT>(pl->coop_C);
    CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (!pl->coop_groups) {
        int per_sm = 0, sms = 0, can = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, kCoopThreads, smem));
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, pl->device);
        cudaDeviceGetAttribute(&can, cudaDevAttrCooperativeLaunch, pl->device);
        long long groups = can ? (long long)per_sm * sms / pl->coop_gsz : 0;
        if (groups > pl->S) groups = pl->S;
        if (groups < 1) return kCoopNoFit;
        const size_t b_bar = align_up((size_t)2 * groups * sizeof(int));
        const size_t b_sum = (size_t)2 * groups * pl->coop_gsz * sizeof(CoopSummary<CarryT>);
        void *mem = nullptr;
        cudaError_t ce = cudaMalloc(&mem, b_bar + b_sum);
        if (ce != cudaSuccess) {
            cudaGetLastError();
            return fail(PYITD_E_NOMEM, std::string("cooperative-path scratch allocation failed: ") + cudaGetErrorString(ce));
        }
        CU(cudaMemsetAsync(mem, 0, b_bar + b_sum, st));       // the kernel leaves the barrier counters at zero
        pl->coop_bar = (int *)mem;
        pl->coop_sum = (char *)mem + b_bar;
        pl->coop_groups = (int)groups;
        pl->ws_bytes += b_bar + b_sum;
    }
    p.bar = pl->coop_bar;
    p.sum = pl->coop_sum;
    p.ngroups = pl->coop_groups;
    void *args[] = {(void *)&p};
    CU(cudaLaunchCooperativeKernel((const void *)k, dim3((unsigned)(pl->coop_groups * pl->coop_gsz)), dim3(kCoopThreads), args,
                                   smem, st));
    return 0;
}
static int run_coop(pyitd_plan *pl, const void *x, void *rotations, void *baselines, int32_t *n_rows,
                    int32_t *knot_counts, int32_t *input_knots, int32_t *stop_kind, int32_t *status, cudaStream_t st) {
    CoopParams p = {};
    p.x = x;
    p.rot = rotations;
    p.bas = baselines;
    p.out_sig_stride = (long long)pl->rows * pl->n;
    p.n_rows = n_rows;
    p.knot_counts = knot_counts;
    p.input_knots = input_knots;
    p.stop_kind = stop_kind;
    p.status = status;
    p.S = (int)pl->S;
    p.n = pl->n;
    p.C = pl->coop_C;
    p.gsz = pl->coop_gsz;
    p.emax = pl->emax;
    p.rows = pl->rows;
    p.min_extrema = pl->min_extrema;
    p.opts = pl->opts;
    pl->launches = 0;
    pl->events_used = 0;
    if (int rc = mark(pl, st)) return rc;
    const bool bas = baselines != nullptr;
    int rc;
    switch (pl->dtype) {
        case PYITD_F64: rc = coop_launch_t<double, double, double>(pl, p, bas, st); break;
        case PYITD_F32_MIXED: rc = coop_launch_t<float, double, float>(pl, p, bas, st); break;
        default: rc = coop_launch_t<float, float, float>(pl, p, bas, st); break;
    }
    if (rc) return rc;
    pl->launches = 1;
    return mark(pl, st);
}

static int pyitd_decompose_device_impl(pyitd_plan *pl, const void *x, void *rotations, void *baselines,
                                      int32_t *n_rows, int32_t *knot_counts, int32_t *input_knots,
                                      int32_t *stop_kind, int32_t *status, void *stream) {
    if (!pl || !x || !rotations || !n_rows || !knot_counts || !status)
        return fail(PYITD_E_INVALID, "null argument");
    if ((pl->opts & kOptBaselines) && !baselines)
        return fail(PYITD_E_INVALID, "plan was created with PYITD_OPT_BASELINES but baselines is null");
    if (!(pl->opts & kOptBaselines)) baselines = nullptr;
    cudaStream_t st = (cudaStream_t)stream;
    if (pl->resident)
        return run_resident(pl, x, rotations, baselines, n_rows, knot_counts, input_knots, stop_kind, status, st);
    if (pl->coop) {
        const int rc = run_coop(pl, x, rotations, baselines, n_rows, knot_counts, input_knots, stop_kind, status, st);
        if (rc != kCoopNoFit) return rc;
        pl->coop = false;                                  // the device cannot hold a group: the look-back kernels take over
    }
    if (int rc = ensure_workspace(pl, st)) return rc;
    pl->launches = 0;
    pl->events_used = 0;
    int *sk = stop_kind ? stop_kind : pl->stop_kind;
    const size_t b_sig = (size_t)pl->S * sizeof(int);
    CU(cudaMemsetAsync(pl->stop_e, 0x7f, b_sig, st));
    CU(cudaMemsetAsync(sk, 0, b_sig, st));
    CU(cudaMemsetAsync(status, 0, b_sig, st));
    CU(cudaMemsetAsync(n_rows, 0, b_sig, st));
    CU(cudaMemsetAsync(knot_counts, 0, b_sig * pl->rows, st));
    if (pl->sweep) return run_sweep(pl, x, rotations, baselines, n_rows, knot_counts, input_knots, sk, status, st);

    // G > 1: fork one launch chain per signal range off the caller's stream and join them at the end; with
    // timing enabled the two events then bracket the whole call instead of every launch
    const int G = effective_groups(pl, x);
    if (G > 1) {
        if (int rc = ensure_group_streams(pl, G)) return rc;
        if (int rc = mark(pl, st)) return rc;
        CU(cudaEventRecord(pl->gfork, st));
        for (int g = 0; g < G; ++g) CU(cudaStreamWaitEvent(pl->gstream[g], pl->gfork, 0));
    }
    auto g_lo = [&](int g) { return pl->S * g / G; };
    for (int g = 0; g < G; ++g) {
        cudaStream_t gs = (G > 1) ? pl->gstream[g] : st;
        if (int rc = run_scan(pl, x, status, input_knots ? input_knots : pl->input_knots, gs, 3, g_lo(g),
                              g_lo(g + 1) - g_lo(g), G == 1))
            return rc;
    }

    // one launch per possible extraction + one trailing fix-up launch; signals that stop early
    // cost an immediate CTA exit, so no host sync is needed to learn the level count
    for (int e = 0; e <= pl->emax + 1; ++e) {
        LevelParams lp;
        lp.in = (e == 0) ? x : pl->carry[(e - 1) & 1];
        lp.carry_out = pl->carry[e & 1];
        lp.fix_src = pl->carry[e & 1];          // X_{e-1} = B_{e-2} lives in carry[(e-2)&1]
        lp.rot = rotations;
        lp.bas = baselines;
        lp.out_sig_stride = (long long)pl->rows * pl->n;
        lp.cur = pl->table[e & 1];
        lp.next = pl->table[(e + 1) & 1];
        lp.desc = pl->desc;
        if (int rc = next_tag(pl, st, &lp.tag)) return rc;
        lp.stop_e = pl->stop_e;
        lp.stop_kind = sk;
        lp.n_rows = n_rows;
        lp.knot_counts = knot_counts;
        lp.status = status;
        lp.n = pl->n;
        lp.tiles = pl->tiles;
        lp.e = e;
        lp.emax = pl->emax;
        lp.rows = pl->rows;
        lp.min_extrema = pl->min_extrema;
        lp.opts = pl->opts;
        // knot baseline pre-pass (one thread per knot) for the signals whose table has become sparse: from level 1 on
        const bool use_ls = pl->ls && e >= 1 && e <= pl->emax && (stream_launchable(pl, lp.in) || strided_launchable(pl, lp.in));
        lp.ls = use_ls ? pl->ls : nullptr;
        lp.lscap = use_ls ? pl->lscap : 0;
        for (int g = 0; g < G; ++g) {
            lp.sig0 = (int)g_lo(g);
            cudaStream_t gs = (G > 1) ? pl->gstream[g] : st;
            const long long nsig = g_lo(g + 1) - g_lo(g);
            if (use_ls) {
                const long long y = (148 * 8 + nsig - 1) / nsig;
                const dim3 grid((unsigned)nsig, (unsigned)(y < 1 ? 1 : y));
                if (pl->carry_elem == 8)
                    knot_ls_kernel<double><<<grid, 256, 0, gs>>>(lp.cur, pl->ls, pl->lscap, lp.sig0, e, pl->stop_e, status);
                else
                    knot_ls_kernel<float><<<grid, 256, 0, gs>>>(lp.cur, pl->ls, pl->lscap, lp.sig0, e, pl->stop_e, status);
                CU(cudaGetLastError());
                pl->launches++;
            }
            CU(launch_level(pl, lp, e == 0, gs, nsig));
            pl->launches++;
        }
        if (strided_launchable(pl, lp.in) && e <= pl->emax) {
            // the strided level kernel leaves flag words + per-tile counts: prefix / stop rule, then the compaction pass
            if (int rc = run_strided_passes(pl, lp, st)) return rc;
        }
        if (G == 1)
            if (int rc = mark(pl, st)) return rc;
    }
    if (G > 1) {
        for (int g = 0; g < G; ++g) {
            CU(cudaEventRecord(pl->gjoin[g], pl->gstream[g]));
            CU(cudaStreamWaitEvent(st, pl->gjoin[g], 0));
        }
        if (int rc = mark(pl, st)) return rc;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// sweep kernel: the whole decomposition of the batch in one persistent launch (itd_sweep.cuh)
// ---------------------------------------------------------------------------------------------
template <typename InT, typename CarryT, typename OutT>
static cudaError_t sweep_launch_t(const SweepParams &p, bool bas, int *grid_cache, cudaStream_t st) {
    constexpr size_t smem = kSweepDynSmem ? sizeof(SweepSmem<CarryT>) : 0;      // static up to 2000 table entries
    auto k = bas ? sweep_kernel<InT, CarryT, OutT, true> : sweep_kernel<InT, CarryT, OutT, false>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    if (smem) {
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    if (*grid_cache == 0) {
        int per_sm = 0, dev = 0, sms = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, kSweepWarps * 32, smem);
        if (e != cudaSuccess) return e;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (per_sm < 1) return cudaErrorLaunchOutOfResources;
        if (const char *env = getenv("PYITD_SWEEP_CTAS_PER_SM")) {       // experiment hook
            const int v = atoi(env);
            if (v >= 1 && v < per_sm) per_sm = v;
        }
        *grid_cache = per_sm * sms;
    }
    long long g = *grid_cache;
    const long long items = p.depth_first ? (long long)p.S : (long long)(p.stage_last - p.stage_first + 1) * p.S;
    if (g > items) g = items;
    k<<<(unsigned)g, kSweepWarps * 32, smem, st>>>(p);
    return cudaGetLastError();
}
static int run_sweep(pyitd_plan *pl, const void *x, void *rotations, void *baselines, int32_t *n_rows,
                     int32_t *knot_counts, int32_t *input_knots, int *sk, int32_t *status, cudaStream_t st) {
    const int stages = pl->emax + 2;                        // the scan + extractions 0 .. emax
    CU(cudaMemsetAsync(pl->sw_ticket, 0, (size_t)(pl->rows + 4) * sizeof(int), st));
    CU(cudaMemsetAsync(pl->sw_done, 0, (size_t)pl->S * 2 * sizeof(int), st));
    SweepParams sp = {};
    sp.x = x;
    sp.carry[0] = pl->carry[0];
    sp.carry[1] = pl->carry[1];
    for (int i = 0; i < 2; ++i) {
        sp.tab[i].tau = pl->table[i].tau;
        sp.tab[i].xk = pl->table[i].xk;
        sp.tab[i].mask = pl->table[i].mask;
        sp.tab[i].rcount = pl->sw_rcount[i];
    }
    sp.rot = rotations;
    sp.bas = baselines;
    sp.out_sig_stride = (long long)pl->rows * pl->n;
    sp.kstride = pl->table[0].kstride;
    sp.mstride = pl->table[0].mstride;
    sp.done = pl->sw_done;
    sp.stats = pl->sw_ticket + pl->rows + 1;               // three counters behind the ticket slots (zeroed with them)
    sp.mid_mask = pl->sw_mid_mask;
    sp.mid_ctas = pl->sw_mid_ctas;
    // PYITD_SWEEP_FUSE=0: every extraction is an item of its own (the round-2 v6 behaviour); thresholds: itd_sweep.cuh
    sp.fuse = pl->sw_mid_ctas > 0 ? 1 : 0;                 // (short signals: decided below, with the item order)
    sp.fuse_min_a = getenv("PYITD_SWEEP_FUSE_MIN_A") ? atoi(getenv("PYITD_SWEEP_FUSE_MIN_A")) : kSweepFuseMinA;
    sp.fuse_max_a = getenv("PYITD_SWEEP_FUSE_MAX_A") ? atoi(getenv("PYITD_SWEEP_FUSE_MAX_A")) : kSweepFuseMaxA;
    sp.fuse_min_b = getenv("PYITD_SWEEP_FUSE_MIN_B") ? atoi(getenv("PYITD_SWEEP_FUSE_MIN_B")) : kSweepFuseMinB;
    if (sp.fuse_min_a <= kSweepProbeKnots) sp.fuse_min_a = kSweepProbeKnots + 1;
    if (sp.fuse_min_b < 1) sp.fuse_min_b = 1;
    if (sp.fuse_max_a > kSweepFuseMaxLimit) sp.fuse_max_a = kSweepFuseMaxLimit;
    sp.stop_e = pl->stop_e;
    sp.stop_kind = sk;
    sp.n_rows = n_rows;
    sp.knot_counts = knot_counts;
    sp.status = status;
    sp.input_knots = input_knots ? input_knots : pl->input_knots;
    sp.stage_ns = nullptr;
    sp.S = (int)pl->S;
    sp.n = pl->n;
    sp.spans = pl->sw_spans;
    sp.spw = pl->sw_spw;
    sp.rs = pl->sw_rs;
    sp.emax = pl->emax;
    sp.rows = pl->rows;
    sp.min_extrema = pl->min_extrema;
    sp.opts = pl->opts;
    // L2 prefetch distance of the sample stream in spans: three short spans ahead on levels with few knots, two of the
    // longer spans of the first levels (a line prefetched too early is evicted again before its span: 39 % extra DRAM
    // reads measured on level 0 at distance 6, profiles/r2/ncu_sweep_v1_metrics.txt)
    sp.pf_sparse = getenv("PYITD_SWEEP_PF_SPARSE") ? atoi(getenv("PYITD_SWEEP_PF_SPARSE")) : 3;
    sp.pf_scan = getenv("PYITD_SWEEP_PF_SCAN") ? atoi(getenv("PYITD_SWEEP_PF_SCAN")) : 4;
    sp.pf_dense = getenv("PYITD_SWEEP_PF_DENSE") ? atoi(getenv("PYITD_SWEEP_PF_DENSE")) : 2;
    sp.pf_fused = getenv("PYITD_SWEEP_PF_FUSED") ? atoi(getenv("PYITD_SWEEP_PF_FUSED")) : 2;
    // short signals: signal-major order keeps each CTA's carry / flags / knot lists in L2 between its stages (all
    // resident CTAs' carries must fit comfortably: 592 CTAs x n x carry bytes <= 48 MB, i.e. n <= ~10 000 fp64 samples)
    sp.depth_first = ((size_t)pl->n * pl->carry_elem * 592 <= ((size_t)48 << 20)) ? 1 : 0;
    if (const char *env = getenv("PYITD_SWEEP_DEPTH")) sp.depth_first = atoi(env) ? 1 : 0;
    // signal-major order keeps the carry in L2 between a signal's extractions: a fused pair saves no DRAM traffic there and its
    // counting pass only adds instructions (config 4: 1.12 -> 1.51 ms with pairs).  PYITD_SWEEP_FUSE=2 forces them on.
    if (sp.depth_first && !(getenv("PYITD_SWEEP_FUSE") && atoi(getenv("PYITD_SWEEP_FUSE")) == 2)) sp.fuse = 0;
    // the prediction alone also picks probe / plain extraction for the items that are not fused (not for short signals either:
    // config 4 1.19 -> 1.40 ms with it, the per-item work there is a few microseconds)
    sp.predict = sp.fuse;
    // PYITD_SWEEP_FUSED_SCAN=1: no scan stage -- extraction 0 finds the knots of the raw input inside its chunk builds and the
    // input's knot lists never exist (-5.7 GB of DRAM traffic per 4096 x 65536 step).  Bit-identical, but level 0 is
    // issue-bound and the extra stencil work costs more than the scan stage it replaces (3.18 ms vs 0.75 + 2.07 ms,
    // profiles/r2/README.md), so the separate stage stays the default.
    // Round 2, later: the variant costs 5000 SASS instructions inside the one kernel every stage runs from, and removing it
    // from the default build made the whole step 0.1 - 0.25 ms faster: it is compiled only with -DPYITD_SWEEP_WITH_FUSED_SCAN
    // (pyitd_has_feature("sweep_fused_scan")), and the switch is ignored otherwise.
#ifdef PYITD_SWEEP_WITH_FUSED_SCAN
    sp.fused_scan = (getenv("PYITD_SWEEP_FUSED_SCAN") && atoi(getenv("PYITD_SWEEP_FUSED_SCAN"))) ? 1 : 0;
#else
    sp.fused_scan = 0;
#endif
    auto launch = [&](int first, int last, int ticket_slot) -> cudaError_t {
        sp.stage_first = first;
        sp.stage_last = last;
        sp.ticket = pl->sw_ticket + ticket_slot;
        const bool bas = baselines != nullptr;
        switch (pl->dtype) {
            case PYITD_F64: return sweep_launch_t<double, double, double>(sp, bas, pl->sw_grid, st);
            case PYITD_F32_MIXED: return sweep_launch_t<float, double, float>(sp, bas, pl->sw_grid, st);
            default: return sweep_launch_t<float, float, float>(sp, bas, pl->sw_grid, st);
        }
    };
    if (int rc = mark(pl, st)) return rc;
    if (pl->timing || getenv("PYITD_SWEEP_PER_STAGE")) {
        // measurement mode: one launch per stage, so that every stage can be timed (and profiled) on its own
        for (int s = 0; s < stages; ++s) {
            if (s > 0 || !sp.fused_scan) {                 // (the event list keeps its scan slot: zero length without one)
                CU(launch(s - 1, s - 1, s));
                pl->launches++;
            }
            if (int rc = mark(pl, st)) return rc;
        }
        return 0;
    }
    CU(launch(sp.fused_scan ? 0 : -1, pl->emax, 0));
    pl->launches = 1;
    return mark(pl, st);
}

extern "C" int pyitd_has_feature(const char *name) {
    if (!name) return 0;
    if (!strcmp(name, "sweep_fused_pairs")) return 1;
#ifdef PYITD_SWEEP_WITH_FUSED_SCAN
    if (!strcmp(name, "sweep_fused_scan")) return 1;
#endif
    return 0;
}

extern "C" int pyitd_plan_sweep_stats(pyitd_plan *pl, int64_t *fused_pairs, int64_t *pairs_skipped, int64_t *pairs_failed) {
    if (!pl || !fused_pairs || !pairs_skipped || !pairs_failed) return fail(PYITD_E_INVALID, "null argument");
    if (!pl->sweep || !pl->ws) return fail(PYITD_E_INVALID, "the plan has not run the sweep kernel");
    DeviceGuard guard(pl->device);
    int v[3] = {0, 0, 0};
    CU(cudaMemcpy(v, pl->sw_ticket + pl->rows + 1, sizeof(v), cudaMemcpyDeviceToHost));      // waits for the device
    *fused_pairs = v[0];
    *pairs_skipped = v[1];
    *pairs_failed = v[2];
    return 0;
}

extern "C" int pyitd_plan_set_groups(pyitd_plan *pl, int groups) {
    if (!pl) return fail(PYITD_E_INVALID, "null plan");
    if (groups < 1 || groups > kMaxGroups) return fail(PYITD_E_INVALID, "groups must be in [1, 16]");
    pl->groups = groups;
    return 0;
}
extern "C" int pyitd_plan_groups(const pyitd_plan *pl) { return pl ? pl->groups : PYITD_E_INVALID; }

// measurement aid: the memory system's ceiling for the level kernel's traffic mix (read 1, write 2 streams).
// chunk_vec == 0: grid-stride (neighbouring blocks touch neighbouring addresses); chunk_vec > 0: block b streams its own
// contiguous range of chunk_vec vectors, then the range of block b + gridDim.x, ... (the one-CTA-per-signal access pattern)
__global__ void __launch_bounds__(256) mix_probe_kernel(const double2 *__restrict__ x, double2 *__restrict__ y,
                                                        double2 *__restrict__ z, long long nvec, long long chunk_vec) {
    if (chunk_vec == 0) {
        const long long stride = (long long)gridDim.x * blockDim.x;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
            const double2 v = x[i];
            y[i] = make_double2(v.x - 1.0, v.y - 1.0);
            z[i] = make_double2(v.x + 1.0, v.y + 1.0);
        }
    } else {
        for (long long c0 = (long long)blockIdx.x * chunk_vec; c0 < nvec; c0 += (long long)gridDim.x * chunk_vec) {
            const long long c1 = (c0 + chunk_vec < nvec) ? c0 + chunk_vec : nvec;
            for (long long i = c0 + threadIdx.x; i < c1; i += blockDim.x) {
                const double2 v = x[i];
                y[i] = make_double2(v.x - 1.0, v.y - 1.0);
                z[i] = make_double2(v.x + 1.0, v.y + 1.0);
            }
        }
    }
}
extern "C" int pyitd_probe_mixed_traffic(const void *x, void *y, void *z, int64_t n_doubles, int ctas, int64_t chunk_doubles,
                                         void *stream) {
    if (!x || !y || !z || n_doubles < 2 || ctas < 1 || chunk_doubles < 0) return fail(PYITD_E_INVALID, "bad argument");
    mix_probe_kernel<<<(unsigned)ctas, 256, 0, (cudaStream_t)stream>>>((const double2 *)x, (double2 *)y, (double2 *)z,
                                                                      n_doubles / 2, chunk_doubles / 2);
    CU(cudaGetLastError());
    return 0;
}

extern "C" int pyitd_plan_enable_timing(pyitd_plan *pl, int enable) {
    if (!pl) return fail(PYITD_E_INVALID, "null plan");
    ON_DEVICE(pl->device);
    if (enable && !pl->events) {
        pl->n_events = pl->emax + 5;
        pl->events = new (std::nothrow) cudaEvent_t[pl->n_events];
        if (!pl->events) return fail(PYITD_E_NOMEM, "host allocation failed");
        for (int i = 0; i < pl->n_events; ++i) CU(cudaEventCreate(&pl->events[i]));
    }
    pl->timing = enable != 0;
    pl->events_used = 0;
    return 0;
}

extern "C" int pyitd_plan_launch_times(pyitd_plan *pl, float *ms, int capacity) {
    if (!pl || !ms) return fail(PYITD_E_INVALID, "null argument");
    if (!pl->timing || pl->events_used < 2) return 0;
    ON_DEVICE(pl->device);
    CU(cudaEventSynchronize(pl->events[pl->events_used - 1]));
    int n = pl->events_used - 1;
    if (n > capacity) n = capacity;
    for (int i = 0; i < n; ++i) CU(cudaEventElapsedTime(&ms[i], pl->events[i], pl->events[i + 1]));
    return n;
}

static int pyitd_extract_level_device_impl(pyitd_plan *pl, const void *x, void *rotation, void *baseline,
                                          int32_t *knot_count, int32_t *status, void *stream) {
    if (!pl || !x || !rotation || !baseline || !knot_count || !status)
        return fail(PYITD_E_INVALID, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = ensure_workspace(pl, st)) return rc;
    pl->launches = 0;
    const size_t b_sig = (size_t)pl->S * sizeof(int);
    CU(cudaMemsetAsync(pl->stop_e, 0x7f, b_sig, st));
    CU(cudaMemsetAsync(status, 0, b_sig, st));
    if (int rc = run_scan(pl, x, status, knot_count, st)) return rc;
    LevelParams lp;
    lp.in = x;
    lp.carry_out = pl->carry[0];
    lp.fix_src = pl->carry[0];
    lp.rot = rotation;
    lp.bas = baseline;
    lp.out_sig_stride = pl->n;
    lp.cur = pl->table[0];
    lp.next = pl->table[1];
    lp.desc = pl->desc;
    if (int rc = next_tag(pl, st, &lp.tag)) return rc;
    lp.stop_e = pl->stop_e;
    lp.stop_kind = pl->stop_kind;
    lp.n_rows = pl->input_knots;            // scratch: the stop bookkeeping is not reported here
    lp.knot_counts = pl->input_knots;
    lp.status = status;
    lp.n = pl->n;
    lp.tiles = pl->tiles;
    lp.e = 0;
    lp.emax = 0x3fffffff;                   // never the "last" level: row 0 is the plain rotation
    lp.rows = 1;
    lp.min_extrema = 0;
    lp.opts = 0;
    CU(launch_level(pl, lp, true, st, pl->S));
    pl->launches++;
    return 0;
}

static int pyitd_extract_with_knots_device_impl(pyitd_plan *pl, const void *x, const int32_t *knots,
                                               int64_t knot_capacity, const int32_t *knot_count, int64_t n_knot_rows,
                                               void *rotation, void *baseline, int32_t *status, void *stream) {
    if (!pl || !x || !knots || !knot_count || !rotation || !baseline || !status || knot_capacity < 0)
        return fail(PYITD_E_INVALID, "null argument");
    if (n_knot_rows != 1 && n_knot_rows != pl->S)
        return fail(PYITD_E_INVALID, "n_knot_rows must be 1 (one list shared by every signal) or n_signals");
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = ensure_workspace(pl, st)) return rc;
    pl->launches = 0;
    const size_t b_sig = (size_t)pl->S * sizeof(int);
    CU(cudaMemsetAsync(pl->stop_e, 0x7f, b_sig, st));
    CU(cudaMemsetAsync(status, 0, b_sig, st));
    const int shared_list = (n_knot_rows == 1 && pl->S != 1) ? 1 : 0;
    switch (pl->dtype) {
        case PYITD_F64:
            table_from_knots_kernel<double, double><<<(unsigned)pl->S, 256, 0, st>>>(
                x, knots, knot_capacity, knot_count, shared_list, pl->table[0], status, pl->n, pl->tiles, pl->tile);
            break;
        case PYITD_F32_MIXED:
            table_from_knots_kernel<float, double><<<(unsigned)pl->S, 256, 0, st>>>(
                x, knots, knot_capacity, knot_count, shared_list, pl->table[0], status, pl->n, pl->tiles, pl->tile);
            break;
        default:
            table_from_knots_kernel<float, float><<<(unsigned)pl->S, 256, 0, st>>>(
                x, knots, knot_capacity, knot_count, shared_list, pl->table[0], status, pl->n, pl->tiles, pl->tile);
    }
    CU(cudaGetLastError());
    pl->launches++;
    LevelParams lp;
    lp.in = x;
    lp.carry_out = pl->carry[0];
    lp.fix_src = pl->carry[0];
    lp.rot = rotation;
    lp.bas = baseline;
    lp.out_sig_stride = pl->n;
    lp.cur = pl->table[0];
    lp.next = pl->table[1];
    lp.desc = pl->desc;
    if (int rc = next_tag(pl, st, &lp.tag)) return rc;
    lp.stop_e = pl->stop_e;
    lp.stop_kind = pl->stop_kind;
    lp.n_rows = pl->input_knots;            // scratch: the stop bookkeeping is not reported here
    lp.knot_counts = pl->input_knots;
    lp.status = status;
    lp.n = pl->n;
    lp.tiles = pl->tiles;
    lp.e = 0;
    lp.emax = 0x3fffffff;                   // never the "last" level: row 0 is the plain rotation
    lp.rows = 1;
    lp.min_extrema = 0;
    lp.opts = kOptGivenKnots;
    CU(launch_level(pl, lp, true, st, pl->S));
    pl->launches++;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// spline-baseline variant: knot scan, truncated-PCR coefficient pass, evaluation pass
// ---------------------------------------------------------------------------------------------
template <typename InT, typename CarryT, typename OutT>
static cudaError_t launch_spline_t(const SplineParams &p, long long S, cudaStream_t st) {
    auto k = spline_level_kernel<InT, CarryT, OutT>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SplineSmem));
    if (e != cudaSuccess) return e;
    k<<<(unsigned)(S * p.parts), kSplThreads, sizeof(SplineSmem), st>>>(p);
    return cudaGetLastError();
}

static int run_spline(pyitd_plan *pl, const void *x, void *rotation, void *baseline, int32_t *knot_count,
                      int32_t *status, int min_knots, cudaStream_t st) {
    if (pl->dtype == PYITD_F32)
        return fail(PYITD_E_INVALID, "the spline variant computes in float64: use a PYITD_F64 or PYITD_F32_MIXED plan");
    if (int rc = ensure_workspace(pl, st)) return rc;
    CU(cudaMemsetAsync(status, 0, (size_t)pl->S * sizeof(int), st));
    if (int rc = run_scan(pl, x, status, knot_count, st)) return rc;
    SplineParams sp;
    sp.x = x;
    sp.rot = rotation;
    sp.bas = baseline;
    sp.tab = pl->table[0];
    sp.status = status;
    sp.n = pl->n;
    sp.tiles = pl->tiles;
    sp.tile = pl->tile;
    sp.min_knots = min_knots < 2 ? 2 : min_knots;
    // one CTA per signal walks its knot windows when there are enough signals to fill the device (4 CTAs per SM),
    // else the windows of a signal are spread over several CTAs
    const long long max_parts = ((long long)pl->n + kSplUseful - 1) / kSplUseful;
    long long parts = (148ll * 8 + pl->S - 1) / pl->S;
    if (parts > max_parts) parts = max_parts;
    if (parts < 1) parts = 1;
    sp.parts = (int)parts;
    // event timing: launch 0 = knot scan, 1 = spline level kernel
    if (pl->dtype == PYITD_F64) CU((launch_spline_t<double, double, double>(sp, pl->S, st)));
    else CU((launch_spline_t<float, double, float>(sp, pl->S, st)));
    pl->launches++;
    return mark(pl, st);
}

static int pyitd_extract_spline_device_impl(pyitd_plan *pl, const void *x, void *rotation, void *baseline,
                                           int32_t *knot_count, int32_t *status, int min_knots, void *stream) {
    if (!pl || !x || !baseline || !knot_count || !status) return fail(PYITD_E_INVALID, "null argument");
    pl->launches = 0;
    pl->events_used = 0;
    return run_spline(pl, x, rotation, baseline, knot_count, status, min_knots, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// 2-D crossways ensemble ITD: row passes, column passes as row passes on the transposed batch
// ---------------------------------------------------------------------------------------------
struct Sift2dShape {
    long long B;
    int H, W;
};
static int sift2d_check(const pyitd_plan *rp, const pyitd_plan *cp, long long B, long long H, long long W) {
    if (!rp || !cp) return fail(PYITD_E_INVALID, "null plan");
    if (B < 1 || H < 3 || W < 3) return fail(PYITD_E_INVALID, "images must be at least 3 x 3");
    if (rp->S != B * H || rp->n != W) return fail(PYITD_E_INVALID, "row plan must be (n_images * height) signals of width samples");
    if (cp->S != B * W || cp->n != H) return fail(PYITD_E_INVALID, "column plan must be (n_images * width) signals of height samples");
    if (rp->dtype != cp->dtype || rp->device != cp->device) return fail(PYITD_E_INVALID, "row and column plans differ in dtype or device");
    if (rp->dtype == PYITD_F32) return fail(PYITD_E_INVALID, "use PYITD_F64 or PYITD_F32_MIXED plans");
    if (B > 65535) return fail(PYITD_E_INVALID, "at most 65535 images per call");
    return 0;
}
template <typename T>
static int sift2d_transpose(const void *in, void *out, Sift2dShape sh, int H, int W, cudaStream_t st) {
    dim3 grid((W + kTrTile - 1) / kTrTile, (H + kTrTile - 1) / kTrTile, (unsigned)sh.B);
    transpose_batch_kernel<T><<<grid, dim3(32, 8), 0, st>>>((const T *)in, (T *)out, H, W);
    CU(cudaGetLastError());
    return 0;
}
// scratch: 3 buffers of B*H*W elements; out may alias none of them
template <typename T>
static int sift2d_crossways(pyitd_plan *rp, pyitd_plan *cp, const void *x, void *out, void *scratch, Sift2dShape sh,
                            int min_knots, cudaStream_t st) {
    const size_t elems = (size_t)sh.B * sh.H * sh.W;
    T *b0 = (T *)scratch, *b1 = b0 + elems, *b2 = b1 + elems;
    if (int rc = ensure_workspace(rp, st)) return rc;
    if (int rc = ensure_workspace(cp, st)) return rc;
    int *rk = rp->input_knots, *rs = rp->stop_kind, *ck = cp->input_knots, *cs = cp->stop_kind;   // per-signal scratch
    // lengthwise[r, :] = base(data[r, :])
    if (int rc = run_spline(rp, x, nullptr, b0, rk, rs, min_knots, st)) return rc;
    // crosswise[:, c] = base(data[:, c])   (kept transposed in b2)
    if (int rc = sift2d_transpose<T>(x, b1, sh, sh.H, sh.W, st)) return rc;
    if (int rc = run_spline(cp, b1, nullptr, b2, ck, cs, min_knots, st)) return rc;
    // crosswise[r, :] = base(crosswise[r, :])
    if (int rc = sift2d_transpose<T>(b2, b1, sh, sh.W, sh.H, st)) return rc;
    if (int rc = run_spline(rp, b1, nullptr, b2, rk, rs, min_knots, st)) return rc;
    // lengthwise[:, c] = base(lengthwise[:, c])   (comes back transposed in b0)
    if (int rc = sift2d_transpose<T>(b0, b1, sh, sh.H, sh.W, st)) return rc;
    if (int rc = run_spline(cp, b1, nullptr, b0, ck, cs, min_knots, st)) return rc;
    // (lengthwise + crosswise) / 2
    dim3 grid((sh.W + kTrTile - 1) / kTrTile, (sh.H + kTrTile - 1) / kTrTile, (unsigned)sh.B);
    transpose_average_kernel<T><<<grid, dim3(32, 8), 0, st>>>(b0, b2, (T *)out, sh.H, sh.W);
    CU(cudaGetLastError());
    rp->launches += 4;
    return 0;
}

extern "C" int64_t pyitd_crossways_scratch_bytes(const pyitd_plan *rp, int64_t n_images, int64_t height, int64_t width) {
    if (!rp) return 0;
    return (int64_t)(3 * (size_t)n_images * (size_t)height * (size_t)width * rp->io_elem);
}

static int pyitd_crossways_device_impl(pyitd_plan *rp, pyitd_plan *cp, const void *images, void *out, void *scratch,
                                      int64_t n_images, int64_t height, int64_t width, int min_knots, void *stream) {
    if (!images || !out || !scratch) return fail(PYITD_E_INVALID, "null argument");
    if (int rc = sift2d_check(rp, cp, n_images, height, width)) return rc;
    rp->launches = cp->launches = 0;
    rp->events_used = cp->events_used = 0;
    const Sift2dShape sh = {n_images, (int)height, (int)width};
    if (rp->dtype == PYITD_F64) return sift2d_crossways<double>(rp, cp, images, out, scratch, sh, min_knots, (cudaStream_t)stream);
    return sift2d_crossways<float>(rp, cp, images, out, scratch, sh, min_knots, (cudaStream_t)stream);
}

extern "C" int64_t pyitd_ensemble2d_scratch_bytes(const pyitd_plan *rp, int64_t draws, int64_t height, int64_t width) {
    if (!rp) return 0;
    return (int64_t)(5 * (size_t)(2 * draws) * (size_t)height * (size_t)width * rp->io_elem);
}

template <typename T>
static int sift2d_ensemble(pyitd_plan *rp, pyitd_plan *cp, const void *image, const void *noise, void *lowpass, void *scratch,
                           int draws, int H, int W, int min_knots, cudaStream_t st) {
    const long long hw = (long long)H * W;
    const size_t elems = (size_t)(2 * draws) * (size_t)hw;
    T *members = (T *)scratch, *result = members + elems, *cw_scratch = result + elems;
    const unsigned grid = (unsigned)((hw + 255) / 256);
    ensemble_members_kernel<T><<<grid, 256, 0, st>>>((const T *)image, (const T *)noise, members, hw, draws);
    CU(cudaGetLastError());
    const Sift2dShape sh = {2ll * draws, H, W};
    if (int rc = sift2d_crossways<T>(rp, cp, members, result, cw_scratch, sh, min_knots, st)) return rc;
    ensemble_mean_kernel<T><<<grid, 256, 0, st>>>(result, (T *)lowpass, hw, draws);
    CU(cudaGetLastError());
    rp->launches += 2;
    return 0;
}

static int pyitd_ensemble2d_device_impl(pyitd_plan *rp, pyitd_plan *cp, const void *image, const void *noise, void *lowpass,
                                       void *scratch, int64_t draws, int64_t height, int64_t width, int min_knots,
                                       void *stream) {
    if (!image || !noise || !lowpass || !scratch) return fail(PYITD_E_INVALID, "null argument");
    if (draws < 1) return fail(PYITD_E_INVALID, "draws must be >= 1");
    if (int rc = sift2d_check(rp, cp, 2 * draws, height, width)) return rc;
    rp->launches = cp->launches = 0;
    rp->events_used = cp->events_used = 0;
    if (rp->dtype == PYITD_F64)
        return sift2d_ensemble<double>(rp, cp, image, noise, lowpass, scratch, (int)draws, (int)height, (int)width, min_knots,
                                       (cudaStream_t)stream);
    return sift2d_ensemble<float>(rp, cp, image, noise, lowpass, scratch, (int)draws, (int)height, (int)width, min_knots,
                                  (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// post-decomposition analytics; no plan: these kernels need no workspace
// ---------------------------------------------------------------------------------------------
extern "C" int pyitd_wpe_device(const void *rows, int64_t n_rows_total, int64_t n_samples, int dtype, int order,
                                int normalize, const int32_t *valid_rows, int64_t rows_per_signal, double *out,
                                void *stream) {
    if (!rows || !out) return fail(PYITD_E_INVALID, "null argument");
    if (order != 3) return fail(PYITD_E_INVALID, "only order 3 is implemented (the only order the reference uses)");
    if (n_rows_total < 1 || n_rows_total > 0x7fffffffll || n_samples < 0 || n_samples > 0x7fffffffll)
        return fail(PYITD_E_INVALID, "bad shape");
    if (valid_rows && rows_per_signal < 1) return fail(PYITD_E_INVALID, "rows_per_signal must be >= 1 with valid_rows");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == PYITD_F64)
        wpe3_kernel<double><<<(unsigned)n_rows_total, 256, 0, st>>>((const double *)rows, n_samples, valid_rows,
                                                                    (int)rows_per_signal, normalize, out);
    else
        wpe3_kernel<float><<<(unsigned)n_rows_total, 256, 0, st>>>((const float *)rows, n_samples, valid_rows,
                                                                   (int)rows_per_signal, normalize, out);
    CU(cudaGetLastError());
    return 0;
}

extern "C" int pyitd_column_fsum_device(const void *rows, int64_t n_signals, int64_t rows_per_signal, int64_t n_samples,
                                        int dtype, const int32_t *valid_rows, double *column_sums, double *totals,
                                        void *stream) {
    if (!rows || !column_sums) return fail(PYITD_E_INVALID, "null argument");
    if (n_signals < 1 || n_signals > 65535 || n_samples < 1) return fail(PYITD_E_INVALID, "bad shape (at most 65535 signals per call)");
    if (rows_per_signal < 1 || rows_per_signal > kFsumMaxRows)
        return fail(PYITD_E_INVALID, "rows_per_signal must be in [1, 64]");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)((n_samples + 255) / 256), (unsigned)n_signals);
    if (dtype == PYITD_F64)
        column_fsum_kernel<double><<<grid, 256, 0, st>>>((const double *)rows, (int)rows_per_signal, n_samples, valid_rows, column_sums);
    else
        column_fsum_kernel<float><<<grid, 256, 0, st>>>((const float *)rows, (int)rows_per_signal, n_samples, valid_rows, column_sums);
    CU(cudaGetLastError());
    if (totals) {
        dd_total_kernel<<<(unsigned)n_signals, 256, 0, st>>>(column_sums, n_samples, totals);
        CU(cudaGetLastError());
    }
    return 0;
}

static int pyitd_find_knots_device_impl(pyitd_plan *pl, const void *x, int kinds, int32_t *knots,
                                       int64_t knot_capacity, int32_t *knot_count, int32_t *status,
                                       void *stream) {
    if (!pl || !x || !knots || !knot_count || !status || knot_capacity < 0)
        return fail(PYITD_E_INVALID, "null argument");
    if (kinds < 1 || kinds > 3) return fail(PYITD_E_INVALID, "kinds must be 1 (valleys), 2 (peaks) or 3 (both)");
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = ensure_workspace(pl, st)) return rc;
    pl->launches = 0;
    CU(cudaMemsetAsync(status, 0, (size_t)pl->S * sizeof(int), st));
    if (int rc = run_scan(pl, x, status, nullptr, st, kinds)) return rc;
    long long per = (knot_capacity + 255) / 256;
    if (per < 1) per = 1;
    if (per > 1024) per = 1024;
    for (long long s0 = 0; s0 < pl->S; s0 += 65535) {
        const long long ns = (pl->S - s0 < 65535) ? pl->S - s0 : 65535;
        dim3 grid((unsigned)per, (unsigned)ns);
        export_knots_kernel<<<grid, 256, 0, st>>>(pl->table[0].tau + s0 * pl->table[0].kstride,
                                                  pl->table[0].kcount + s0, pl->table[0].kstride,
                                                  knots + s0 * knot_capacity, knot_capacity, knot_count + s0);
        CU(cudaGetLastError());
        pl->launches++;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// host-buffer entry point: the batch is walked in chunks through two slots so that the H2D copy of
// chunk c+1 and its kernels overlap the D2H copy of chunk c (PCIe is the bottleneck of this call:
// ~8 x rows bytes come back for every 8 bytes that go in).  Only the rows a signal actually
// produced are copied back unless PYITD_OPT_ZERO_TAIL asks for the zero-filled tail.
// ---------------------------------------------------------------------------------------------
// bytes one host-pipeline slot needs per signal: the device mirrors of x / rotations / baselines plus the sub-plan's
// workspace (two carry buffers, two knot tables, flag masks; the knot_ls table of the stream / strided paths)
static size_t host_slot_bytes_per_signal(const pyitd_plan *pl) {
    const size_t n = (size_t)pl->n;
    const size_t io = n * pl->io_elem, out = io * pl->rows;
    const size_t ws = 2 * n * pl->carry_elem + 2 * (n + 8) * (sizeof(int) + pl->carry_elem) + n / 2 + n * pl->carry_elem;
    return io + out * ((pl->opts & kOptBaselines) ? 2 : 1) + ws;
}

static long long pick_host_chunk(const pyitd_plan *pl) {
    if (const char *env = getenv("PYITD_HOST_CHUNK")) {
        long long v = atoll(env);
        if (v >= 1) return v < pl->S ? v : pl->S;
    }
    // ~128 MiB of input per chunk, at least 4 chunks for the pipeline to overlap anything, and -- when the signals are
    // short enough for the one-CTA-per-signal kernels -- at least 256 signals (those kernels want a full GPU)
    long long c = (long long)((128ull << 20) / ((size_t)pl->n * pl->io_elem));
    const bool stream_shape = (pl->n % 4 == 0) && (((long long)pl->n + kStreamTile - 1) / kStreamTile <= kStreamMaxTiles);
    if (c < 256 && stream_shape) c = 256;
    if (c > pl->S / 4 && pl->S >= 1024) c = pl->S / 4;
    // two slots must fit in 70 % of the free device memory: long signals get small chunks instead of E_NOMEM
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
        const size_t per = host_slot_bytes_per_signal(pl);
        const long long cap = (long long)((double)free_b * 0.7 / 2.0 / (double)per);
        if (c > cap) c = cap;
    } else {
        cudaGetLastError();
    }
    if (c > pl->S) c = pl->S;
    return c < 1 ? 1 : c;
}

static void free_host_slots(pyitd_plan *pl) {
    for (auto &sl : pl->slot) {
        if (sl.sub) pyitd_plan_destroy(sl.sub);
        cudaFree(sl.x);
        cudaFree(sl.rot);
        cudaFree(sl.bas);
        cudaFree(sl.ints);
        if (sl.h_rows) cudaFreeHost(sl.h_rows);
        if (sl.stream) cudaStreamDestroy(sl.stream);
        sl = pyitd_plan::HostSlot();
    }
    pl->host_chunk = 0;
}

static int create_host_slots(pyitd_plan *pl, long long C) {
    const bool want_bas = (pl->opts & kOptBaselines) != 0;
    const int nslots = (C < pl->S) ? 2 : 1;
    for (int b = 0; b < nslots; ++b) {
        auto &sl = pl->slot[b];
        if (int rc = pyitd_plan_create(&sl.sub, pl->device, C, pl->n, pl->dtype, pl->max_iteration,
                                       pl->min_extrema, (int)pl->opts))
            return rc;
        const size_t b_in = (size_t)C * pl->n * pl->io_elem, b_out = b_in * pl->rows;
        CU(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
        CU(cudaMalloc(&sl.x, b_in));
        CU(cudaMalloc(&sl.rot, b_out));
        if (want_bas) CU(cudaMalloc(&sl.bas, b_out));
        CU(cudaMalloc((void **)&sl.ints, (4 + (size_t)pl->rows) * (size_t)C * sizeof(int)));
        CU(cudaMallocHost((void **)&sl.h_rows, 2 * (size_t)C * sizeof(int)));
    }
    return 0;
}

static int ensure_host_slots(pyitd_plan *pl) {
    if (pl->host_chunk) return 0;
    const long long C = pick_host_chunk(pl);
    if (int rc = create_host_slots(pl, C)) {
        // a slot that was only partly created must not survive: the next call would overwrite (and leak) it
        const std::string msg = g_err;
        cudaGetLastError();
        free_host_slots(pl);
        return fail(rc, msg);
    }
    pl->host_chunk = C;
    return 0;
}

// the chunk loop of pyitd_decompose_host; on an error the caller drains both slot streams before returning
static int host_chunk_loop(pyitd_plan *pl, const void *x, void *rotations, void *baselines,
                           int32_t *n_rows, int32_t *knot_counts, int32_t *input_knots,
                           int32_t *stop_kind, int32_t *status) {
    const bool want_bas = (pl->opts & kOptBaselines) != 0;
    const long long C = pl->host_chunk;
    const size_t row_b = (size_t)pl->n * pl->io_elem;          // bytes of one row
    const size_t sig_out_b = row_b * pl->rows;                 // bytes of one signal's output block
    const bool all_rows = (pl->opts & kOptZeroTail) != 0;
    pl->launches = 0;
    int nchunk = 0;
    for (long long s0 = 0; s0 < pl->S; s0 += C, ++nchunk) {
        auto &sl = pl->slot[nchunk & 1];
        const long long cs = (pl->S - s0 < C) ? pl->S - s0 : C;
        cudaStream_t st = sl.stream;
        // the slot's previous chunk (two chunks ago) must have left the device buffers
        CU(cudaStreamSynchronize(st));
        int *d_nrows = sl.ints, *d_counts = d_nrows + C, *d_ik = d_counts + C * pl->rows, *d_kind = d_ik + C,
            *d_status = d_kind + C;
        CU(cudaMemcpyAsync(sl.x, (const char *)x + (size_t)s0 * row_b, (size_t)cs * row_b, cudaMemcpyHostToDevice, st));
        // a short last chunk runs on the full-size sub-plan: its tail signals decompose zeros (monotone input: one
        // level, one all-zero row) instead of whatever the slot held, and are never copied back
        if (cs < C) CU(cudaMemsetAsync((char *)sl.x + (size_t)cs * row_b, 0, (size_t)(C - cs) * row_b, st));
        if (int rc = pyitd_decompose_device(sl.sub, sl.x, sl.rot, want_bas ? sl.bas : nullptr, d_nrows, d_counts,
                                            d_ik, d_kind, d_status, st))
            return rc;
        pl->launches += sl.sub->launches;
        CU(cudaMemcpyAsync(sl.h_rows, d_nrows, (size_t)cs * sizeof(int), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(sl.h_rows + C, d_kind, (size_t)cs * sizeof(int), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(knot_counts + s0 * pl->rows, d_counts, (size_t)cs * pl->rows * sizeof(int),
                           cudaMemcpyDeviceToHost, st));
        if (input_knots) CU(cudaMemcpyAsync(input_knots + s0, d_ik, (size_t)cs * sizeof(int), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(status + s0, d_status, (size_t)cs * sizeof(int), cudaMemcpyDeviceToHost, st));
        if (all_rows) {
            CU(cudaMemcpyAsync((char *)rotations + (size_t)s0 * sig_out_b, sl.rot, (size_t)cs * sig_out_b,
                               cudaMemcpyDeviceToHost, st));
            if (want_bas)
                CU(cudaMemcpyAsync((char *)baselines + (size_t)s0 * sig_out_b, sl.bas, (size_t)cs * sig_out_b,
                                   cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
        } else {
            // the kernels of this chunk are done once n_rows is on the host; the other slot's D2H
            // keeps the copy engine busy while this thread waits here
            CU(cudaStreamSynchronize(st));
            for (long long s = 0; s < cs; ++s) {
                int nr = sl.h_rows[s];
                if (nr < 0) nr = 0;
                if (nr > pl->rows) nr = pl->rows;
                if (nr == 0) continue;
                CU(cudaMemcpyAsync((char *)rotations + (size_t)(s0 + s) * sig_out_b, (const char *)sl.rot + (size_t)s * sig_out_b,
                                   (size_t)nr * row_b, cudaMemcpyDeviceToHost, st));
                if (want_bas) {
                    // valid baseline rows: n_rows on the iteration stop, n_rows - 1 on the knot stop
                    const int nb = (sl.h_rows[C + s] == kStopIter) ? nr : nr - 1;
                    if (nb > 0)
                        CU(cudaMemcpyAsync((char *)baselines + (size_t)(s0 + s) * sig_out_b,
                                           (const char *)sl.bas + (size_t)s * sig_out_b, (size_t)nb * row_b,
                                           cudaMemcpyDeviceToHost, st));
                }
            }
        }
        for (long long s = 0; s < cs; ++s) n_rows[s0 + s] = sl.h_rows[s];
        if (stop_kind)
            for (long long s = 0; s < cs; ++s) stop_kind[s0 + s] = sl.h_rows[C + s];
    }
    return 0;
}

static int pyitd_decompose_host_impl(pyitd_plan *pl, const void *x, void *rotations, void *baselines,
                                    int32_t *n_rows, int32_t *knot_counts, int32_t *input_knots,
                                    int32_t *stop_kind, int32_t *status) {
    if (!pl || !x || !rotations || !n_rows || !knot_counts || !status)
        return fail(PYITD_E_INVALID, "null argument");
    if ((pl->opts & kOptBaselines) && !baselines) return fail(PYITD_E_INVALID, "baselines is null");
    if (int rc = ensure_host_slots(pl)) return rc;
    const int rc = host_chunk_loop(pl, x, rotations, baselines, n_rows, knot_counts, input_knots, stop_kind, status);
    // success or not, nothing of this call may still be in flight when it returns: the other slot's copies write into
    // the caller's buffers and its kernels use the slot's workspace
    const std::string msg = g_err;
    cudaError_t de = cudaSuccess;
    for (auto &sl : pl->slot)
        if (sl.stream) {
            const cudaError_t e = cudaStreamSynchronize(sl.stream);
            if (e != cudaSuccess && de == cudaSuccess) de = e;
        }
    if (rc) return fail(rc, msg);
    if (de != cudaSuccess) return fail(PYITD_E_CUDA, std::string("cudaStreamSynchronize: ") + cudaGetErrorString(de));
    return 0;
}

// ---------------------------------------------------------------------------------------------
// public entry points: one call at a time per plan (host lock), on the plan's device (restored on return), ordered
// after the plan's previous call when that ran on another stream
// ---------------------------------------------------------------------------------------------
extern "C" int pyitd_decompose_device(pyitd_plan *pl, const void *x, void *rotations, void *baselines, int32_t *n_rows, int32_t *knot_counts, int32_t *input_knots, int32_t *stop_kind, int32_t *status, void *stream) {
    if (!pl) return fail(PYITD_E_INVALID, "null plan");
    std::lock_guard<std::mutex> lk(pl->mu);
    ON_DEVICE(pl->device);
    if (int rc = plan_acquire(pl, (cudaStream_t)stream)) return rc;
    const int rc = pyitd_decompose_device_impl(pl, x, rotations, baselines, n_rows, knot_counts, input_knots, stop_kind, status, stream);
    if (rc == 0) {
        if (int rr = plan_release(pl, (cudaStream_t)stream)) return rr;
    }
    return rc;
}

extern "C" int pyitd_extract_level_device(pyitd_plan *pl, const void *x, void *rotation, void *baseline, int32_t *knot_count, int32_t *status, void *stream) {
    if (!pl) return fail(PYITD_E_INVALID, "null plan");
    std::lock_guard<std::mutex> lk(pl->mu);
    ON_DEVICE(pl->device);
    if (int rc = plan_acquire(pl, (cudaStream_t)stream)) return rc;
    const int rc = pyitd_extract_level_device_impl(pl, x, rotation, baseline, knot_count, status, stream);
    if (rc == 0) {
        if (int rr = plan_release(pl, (cudaStream_t)stream)) return rr;
    }
    return rc;
}

extern "C" int pyitd_extract_with_knots_device(pyitd_plan *pl, const void *x, const int32_t *knots, int64_t knot_capacity, const int32_t *knot_count, int64_t n_knot_rows, void *rotation, void *baseline, int32_t *status, void *stream) {
    if (!pl) return fail(PYITD_E_INVALID, "null plan");
    std::lock_guard<std::mutex> lk(pl->mu);
    ON_DEVICE(pl->device);
    if (int rc = plan_acquire(pl, (cudaStream_t)stream)) return rc;
    const int rc = pyitd_extract_with_knots_device_impl(pl, x, knots, knot_capacity, knot_count, n_knot_rows, rotation, baseline, status, stream);
    if (rc == 0) {
        if (int rr = plan_release(pl, (cudaStream_t)stream)) return rr;
    }
    return rc;
}

extern "C" int pyitd_extract_spline_device(pyitd_plan *pl, const void *x, void *rotation, void *baseline, int32_t *knot_count, int32_t *status, int min_knots, void *stream) {
    if (!pl) return fail(PYITD_E_INVALID, "null plan");
    std::lock_guard<std::mutex> lk(pl->mu);
    ON_DEVICE(pl->device);
    if (int rc = plan_acquire(pl, (cudaStream_t)stream)) return rc;
    const int rc = pyitd_extract_spline_device_impl(pl, x, rotation, baseline, knot_count, status, min_knots, stream);
    if (rc == 0) {
        if (int rr = plan_release(pl, (cudaStream_t)stream)) return rr;
    }
    return rc;
}

extern "C" int pyitd_find_knots_device(pyitd_plan *pl, const void *x, int kinds, int32_t *knots, int64_t knot_capacity, int32_t *knot_count, int32_t *status, void *stream) {
    if (!pl) return fail(PYITD_E_INVALID, "null plan");
    std::lock_guard<std::mutex> lk(pl->mu);
    ON_DEVICE(pl->device);
    if (int rc = plan_acquire(pl, (cudaStream_t)stream)) return rc;
    const int rc = pyitd_find_knots_device_impl(pl, x, kinds, knots, knot_capacity, knot_count, status, stream);
    if (rc == 0) {
        if (int rr = plan_release(pl, (cudaStream_t)stream)) return rr;
    }
    return rc;
}

extern "C" int pyitd_crossways_device(pyitd_plan *rp, pyitd_plan *cp, const void *images, void *out, void *scratch, int64_t n_images, int64_t height, int64_t width, int min_knots, void *stream) {
    if (!rp || !cp) return fail(PYITD_E_INVALID, "null plan");
    std::unique_lock<std::mutex> lk0(rp->mu, std::defer_lock), lk1(cp->mu, std::defer_lock);
    if (rp == cp) lk0.lock(); else std::lock(lk0, lk1);
    ON_DEVICE(rp->device);
    if (int rc = plan_acquire(rp, (cudaStream_t)stream)) return rc;
    if (int rc = plan_acquire(cp, (cudaStream_t)stream)) return rc;
    const int rc = pyitd_crossways_device_impl(rp, cp, images, out, scratch, n_images, height, width, min_knots, stream);
    if (rc == 0) {
        if (int rr = plan_release(rp, (cudaStream_t)stream)) return rr;
        if (int rr = plan_release(cp, (cudaStream_t)stream)) return rr;
    }
    return rc;
}

extern "C" int pyitd_ensemble2d_device(pyitd_plan *rp, pyitd_plan *cp, const void *image, const void *noise, void *lowpass, void *scratch, int64_t draws, int64_t height, int64_t width, int min_knots, void *stream) {
    if (!rp || !cp) return fail(PYITD_E_INVALID, "null plan");
    std::unique_lock<std::mutex> lk0(rp->mu, std::defer_lock), lk1(cp->mu, std::defer_lock);
    if (rp == cp) lk0.lock(); else std::lock(lk0, lk1);
    ON_DEVICE(rp->device);
    if (int rc = plan_acquire(rp, (cudaStream_t)stream)) return rc;
    if (int rc = plan_acquire(cp, (cudaStream_t)stream)) return rc;
    const int rc = pyitd_ensemble2d_device_impl(rp, cp, image, noise, lowpass, scratch, draws, height, width, min_knots, stream);
    if (rc == 0) {
        if (int rr = plan_release(rp, (cudaStream_t)stream)) return rr;
        if (int rr = plan_release(cp, (cudaStream_t)stream)) return rr;
    }
    return rc;
}

extern "C" int pyitd_decompose_host(pyitd_plan *pl, const void *x, void *rotations, void *baselines, int32_t *n_rows, int32_t *knot_counts, int32_t *input_knots, int32_t *stop_kind, int32_t *status) {
    if (!pl) return fail(PYITD_E_INVALID, "null plan");
    std::lock_guard<std::mutex> lk(pl->mu);
    ON_DEVICE(pl->device);
    const int rc = pyitd_decompose_host_impl(pl, x, rotations, baselines, n_rows, knot_counts, input_knots, stop_kind, status);
    return rc;
}

