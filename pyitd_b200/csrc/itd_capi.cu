// itd_capi.cu -- the extern "C" boundary declared in include/pyitd_b200.h: plan/workspace
// management and the launch sequence of one decomposition.  No torch types, no host syncs on the
// device entry points.
#include "../../include/pyitd_b200.h"

#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>

#include "itd_kernels.cuh"
#include "itd_stream.cuh"

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
    size_t carry_elem = 8, io_elem = 8;
    // workspace
    void *ws = nullptr;
    size_t ws_bytes = 0;
    void *carry[2] = {nullptr, nullptr};
    KnotTable table[2];
    unsigned long long *desc = nullptr;
    int *stop_e = nullptr, *stop_kind = nullptr, *input_knots = nullptr;
    unsigned tag = 0;
    int launches = 0;
    // optional per-launch CUDA-event timing (bench.py's roofline leg)
    bool timing = false;
    cudaEvent_t *events = nullptr;
    int n_events = 0, events_used = 0;
    // lazily allocated device mirrors for the _host entry point
    void *h_x = nullptr, *h_rot = nullptr, *h_bas = nullptr;
    int *h_ints = nullptr;    // n_rows | knot_counts | input_knots | stop_kind | status
    cudaStream_t h_stream = nullptr;
};

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
constexpr int kStreamWarps = 8, kStreamItems = 4, kStreamStages = 2;
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

static cudaError_t launch_scan(const pyitd_plan *pl, const ScanParams &p, cudaStream_t st) {
    if (pl->stream && (reinterpret_cast<uintptr_t>(p.x) & 15) == 0) {
        switch (pl->dtype) {
            case PYITD_F64: return launch_scan_stream_t<double, double>(p, pl->S, st);
            case PYITD_F32_MIXED: return launch_scan_stream_t<float, double>(p, pl->S, st);
            default: return launch_scan_stream_t<float, float>(p, pl->S, st);
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
static cudaError_t launch_level(const pyitd_plan *pl, const LevelParams &p, bool first, cudaStream_t st) {
    if (pl->stream && (reinterpret_cast<uintptr_t>(p.in) & 15) == 0) {
        switch (pl->dtype) {
            case PYITD_F64: return launch_stream_t<double, double, double>(p, pl->S, st);
            case PYITD_F32_MIXED:
                return first ? launch_stream_t<float, double, float>(p, pl->S, st)
                             : launch_stream_t<double, double, float>(p, pl->S, st);
            default: return launch_stream_t<float, float, float>(p, pl->S, st);
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
    CU(cudaSetDevice(device));
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

    int cfg = (n_samples <= 512) ? 0 : 1;
    if (const char *env = getenv("PYITD_TILE_CFG")) {
        int v = atoi(env);
        if (v >= 0 && v < kNumTileCfgs) cfg = v;
    }
    // path choice: many signals -> one pipelined CTA per signal; few long signals -> multi-CTA
    // look-back tiles.  The streaming kernel needs 16-byte aligned rows for its TMA bulk copies.
    const long long stream_tiles = (n_samples + kStreamTile - 1) / kStreamTile;
    bool stream = (n_samples % 4 == 0) && stream_tiles <= kStreamMaxTiles && n_signals >= 256;
    if (const char *env = getenv("PYITD_FORCE_PATH")) {
        if (!strcmp(env, "stream")) stream = (n_samples % 4 == 0) && stream_tiles <= kStreamMaxTiles;
        if (!strcmp(env, "lookback")) stream = false;
    }
    if (stream) cfg = 1;                       // both kernels must agree on the 1024-sample tile
    pl->stream = stream;
    pl->tile_cfg = cfg;
    pl->tile = kTileCfgs[cfg].threads * kTileCfgs[cfg].items;
    pl->tiles = (int)((n_samples + pl->tile - 1) / pl->tile);
    if ((long long)pl->tiles * pl->S > 0x7fffffffll) {
        delete pl;
        return fail(PYITD_E_INVALID, "n_signals * tiles exceeds the grid limit; split the batch");
    }

    const size_t SN = (size_t)pl->S * (size_t)pl->n;
    const long long kstride = (((long long)pl->n + 3) & ~3ll) + 4;
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
    size_t total = 2 * b_carry + 2 * (b_tau + b_xk + b_tbase + b_sig + b_endl + b_mask) + b_desc + 3 * b_sig;
    pl->ws_bytes = total;
    cudaError_t ce = cudaMalloc(&pl->ws, total);
    if (ce != cudaSuccess) {
        cudaGetLastError();
        delete pl;
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
    ce = cudaMemset(pl->desc, 0, b_desc);
    if (ce != cudaSuccess) {
        cudaFree(pl->ws);
        delete pl;
        return fail(PYITD_E_CUDA, std::string("cudaMemset: ") + cudaGetErrorString(ce));
    }
    *out = pl;
    return 0;
}

extern "C" void pyitd_plan_destroy(pyitd_plan *pl) {
    if (!pl) return;
    cudaSetDevice(pl->device);
    if (pl->h_stream) cudaStreamDestroy(pl->h_stream);
    if (pl->events) {
        for (int i = 0; i < pl->n_events; ++i) cudaEventDestroy(pl->events[i]);
        delete[] pl->events;
    }
    cudaFree(pl->h_x);
    cudaFree(pl->h_rot);
    cudaFree(pl->h_bas);
    cudaFree(pl->h_ints);
    cudaFree(pl->ws);
    delete pl;
}

extern "C" int pyitd_plan_rows(const pyitd_plan *pl) { return pl ? pl->rows : PYITD_E_INVALID; }
extern "C" int64_t pyitd_plan_workspace_bytes(const pyitd_plan *pl) { return pl ? (int64_t)pl->ws_bytes : 0; }
extern "C" int pyitd_plan_launches(const pyitd_plan *pl) { return pl ? pl->launches : 0; }

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

static int run_scan(pyitd_plan *pl, const void *x, int *status, int *input_knots, cudaStream_t st, int kinds = 3) {
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
    if (int rc = mark(pl, st)) return rc;
    CU(launch_scan(pl, sp, st));
    pl->launches++;
    return mark(pl, st);
}

extern "C" int pyitd_decompose_device(pyitd_plan *pl, const void *x, void *rotations, void *baselines,
                                      int32_t *n_rows, int32_t *knot_counts, int32_t *input_knots,
                                      int32_t *stop_kind, int32_t *status, void *stream) {
    if (!pl || !x || !rotations || !n_rows || !knot_counts || !status)
        return fail(PYITD_E_INVALID, "null argument");
    if ((pl->opts & kOptBaselines) && !baselines)
        return fail(PYITD_E_INVALID, "plan was created with PYITD_OPT_BASELINES but baselines is null");
    if (!(pl->opts & kOptBaselines)) baselines = nullptr;
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(pl->device));
    pl->launches = 0;
    pl->events_used = 0;
    int *sk = stop_kind ? stop_kind : pl->stop_kind;
    const size_t b_sig = (size_t)pl->S * sizeof(int);
    CU(cudaMemsetAsync(pl->stop_e, 0x7f, b_sig, st));
    CU(cudaMemsetAsync(sk, 0, b_sig, st));
    CU(cudaMemsetAsync(status, 0, b_sig, st));
    CU(cudaMemsetAsync(n_rows, 0, b_sig, st));
    CU(cudaMemsetAsync(knot_counts, 0, b_sig * pl->rows, st));

    if (int rc = run_scan(pl, x, status, input_knots ? input_knots : pl->input_knots, st)) return rc;

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
        CU(launch_level(pl, lp, e == 0, st));
        pl->launches++;
        if (int rc = mark(pl, st)) return rc;
    }
    return 0;
}

extern "C" int pyitd_plan_enable_timing(pyitd_plan *pl, int enable) {
    if (!pl) return fail(PYITD_E_INVALID, "null plan");
    CU(cudaSetDevice(pl->device));
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
    CU(cudaSetDevice(pl->device));
    CU(cudaEventSynchronize(pl->events[pl->events_used - 1]));
    int n = pl->events_used - 1;
    if (n > capacity) n = capacity;
    for (int i = 0; i < n; ++i) CU(cudaEventElapsedTime(&ms[i], pl->events[i], pl->events[i + 1]));
    return n;
}

extern "C" int pyitd_extract_level_device(pyitd_plan *pl, const void *x, void *rotation, void *baseline,
                                          int32_t *knot_count, int32_t *status, void *stream) {
    if (!pl || !x || !rotation || !baseline || !knot_count || !status)
        return fail(PYITD_E_INVALID, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(pl->device));
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
    CU(launch_level(pl, lp, true, st));
    pl->launches++;
    return 0;
}

extern "C" int pyitd_find_knots_device(pyitd_plan *pl, const void *x, int kinds, int32_t *knots,
                                       int64_t knot_capacity, int32_t *knot_count, int32_t *status,
                                       void *stream) {
    if (!pl || !x || !knots || !knot_count || !status || knot_capacity < 0)
        return fail(PYITD_E_INVALID, "null argument");
    if (kinds < 1 || kinds > 3) return fail(PYITD_E_INVALID, "kinds must be 1 (valleys), 2 (peaks) or 3 (both)");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(pl->device));
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
// host-buffer entry point: H2D, decompose, D2H, synchronise
// ---------------------------------------------------------------------------------------------
extern "C" int pyitd_decompose_host(pyitd_plan *pl, const void *x, void *rotations, void *baselines,
                                    int32_t *n_rows, int32_t *knot_counts, int32_t *input_knots,
                                    int32_t *stop_kind, int32_t *status) {
    if (!pl || !x || !rotations || !n_rows || !knot_counts || !status)
        return fail(PYITD_E_INVALID, "null argument");
    const bool want_bas = (pl->opts & kOptBaselines) != 0;
    if (want_bas && !baselines) return fail(PYITD_E_INVALID, "baselines is null");
    CU(cudaSetDevice(pl->device));
    const size_t SN = (size_t)pl->S * pl->n;
    const size_t b_in = SN * pl->io_elem, b_out = b_in * pl->rows;
    const size_t S = (size_t)pl->S;
    if (!pl->h_stream) CU(cudaStreamCreateWithFlags(&pl->h_stream, cudaStreamNonBlocking));
    if (!pl->h_x) CU(cudaMalloc(&pl->h_x, b_in));
    if (!pl->h_rot) CU(cudaMalloc(&pl->h_rot, b_out));
    if (want_bas && !pl->h_bas) CU(cudaMalloc(&pl->h_bas, b_out));
    if (!pl->h_ints) CU(cudaMalloc((void **)&pl->h_ints, (4 + (size_t)pl->rows) * S * sizeof(int)));
    int *d_nrows = pl->h_ints, *d_counts = d_nrows + S, *d_ik = d_counts + S * pl->rows,
        *d_kind = d_ik + S, *d_status = d_kind + S;
    cudaStream_t st = pl->h_stream;
    CU(cudaMemcpyAsync(pl->h_x, x, b_in, cudaMemcpyHostToDevice, st));
    if (int rc = pyitd_decompose_device(pl, pl->h_x, pl->h_rot, want_bas ? pl->h_bas : nullptr, d_nrows,
                                        d_counts, d_ik, d_kind, d_status, st))
        return rc;
    CU(cudaMemcpyAsync(rotations, pl->h_rot, b_out, cudaMemcpyDeviceToHost, st));
    if (want_bas) CU(cudaMemcpyAsync(baselines, pl->h_bas, b_out, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(n_rows, d_nrows, S * sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(knot_counts, d_counts, S * pl->rows * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (input_knots) CU(cudaMemcpyAsync(input_knots, d_ik, S * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (stop_kind) CU(cudaMemcpyAsync(stop_kind, d_kind, S * sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(status, d_status, S * sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return 0;
}
