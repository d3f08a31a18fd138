// torch_ops.cpp -- the thin PyTorch extension over the C ABI ("Native: a TORCH_LIBRARY op set").
//
//   torch.ops.pyitd.find_knots(x, kinds)        replaces detect_peaks + the knot merge     (ITD.py:33-76, :87-98)
//   torch.ops.pyitd.extract_level(x)            replaces itd_baseline_extract              (ITD.py:79-121)
//   torch.ops.pyitd.decompose(x, ...)           replaces ITD.itd for a batch of signals    (ITD.py:351-433)
//
// This file holds NO kernel and NO arithmetic: every op allocates its outputs with torch's caching allocator, takes
// torch's current CUDA stream for the input's device and calls libpyitd_b200.so (include/pyitd_b200.h).  Only the CUDA
// dispatch key is implemented (plus Meta for shape inference): a CPU tensor raises torch's own "could not run ... with
// arguments from the 'CPU' backend" -- there is no CPU fallback.  Built by g++ (no device code), linked against
// libpyitd_b200.so through $ORIGIN.
#include <ATen/ATen.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/library.h>

#include <list>
#include <memory>
#include <mutex>
#include <string>
#include <tuple>

#include "../../include/pyitd_b200.h"

namespace {

using at::Tensor;

// ---- plan cache: a plan owns the HBM workspace of one batch shape (GBs for big batches) ---------------------------
using PlanKey = std::tuple<int, int64_t, int64_t, int, int, int, int>;
// a caller holds a reference for the duration of its call: an LRU eviction from another thread only drops the cache's
// reference, the plan is destroyed when its last user returns (the C library itself serialises calls on one plan)
using PlanRef = std::shared_ptr<pyitd_plan>;
struct PlanEntry { PlanKey key; PlanRef plan; };
std::mutex g_mu;
std::list<PlanEntry> g_plans;               // most recently used first
constexpr size_t kMaxPlans = 4;

PlanRef get_plan(int device, int64_t S, int64_t N, int dtype, int max_iteration, int min_extrema, int options) {
    PlanKey key{device, S, N, dtype, max_iteration, min_extrema, options};
    std::lock_guard<std::mutex> lock(g_mu);
    for (auto it = g_plans.begin(); it != g_plans.end(); ++it)
        if (it->key == key) {
            g_plans.splice(g_plans.begin(), g_plans, it);
            return g_plans.front().plan;
        }
    while (g_plans.size() >= kMaxPlans) g_plans.pop_back();
    pyitd_plan *p = nullptr;
    int rc = pyitd_plan_create(&p, device, S, N, dtype, max_iteration, min_extrema, options);
    TORCH_CHECK(rc == 0, "pyitd_plan_create failed (", rc, "): ", pyitd_last_error());
    g_plans.push_front({key, PlanRef(p, [](pyitd_plan *q) { pyitd_plan_destroy(q); })});
    return g_plans.front().plan;
}

void clear_plans() {
    std::lock_guard<std::mutex> lock(g_mu);
    g_plans.clear();
}

// ---- argument handling --------------------------------------------------------------------------------------------
int precision_code(const Tensor &x, const std::string &precision) {
    if (precision.empty() || precision == "auto") {
        TORCH_CHECK(x.scalar_type() == at::kDouble || x.scalar_type() == at::kFloat,
                    "pyitd: expected a float64 or float32 tensor, got ", x.scalar_type());
        return x.scalar_type() == at::kDouble ? PYITD_F64 : PYITD_F32_MIXED;
    }
    if (precision == "f64") { TORCH_CHECK(x.scalar_type() == at::kDouble, "pyitd: precision f64 needs float64 input"); return PYITD_F64; }
    TORCH_CHECK(x.scalar_type() == at::kFloat, "pyitd: precision ", precision, " needs float32 input");
    if (precision == "f32_mixed") return PYITD_F32_MIXED;
    if (precision == "f32") return PYITD_F32;
    TORCH_CHECK(false, "pyitd: precision must be auto, f64, f32_mixed or f32, got ", precision);
}

Tensor as_batch(const Tensor &x) {
    TORCH_CHECK(x.dim() == 1 || x.dim() == 2, "pyitd: expected a 1-D signal or a 2-D [signals, samples] batch, got ", x.dim(), " dims");
    Tensor b = x.dim() == 1 ? x.unsqueeze(0) : x;
    TORCH_CHECK(b.size(1) >= 3, "pyitd: signal shorter than 3 samples (undefined in the reference, ITD.py:42-43)");
    return b.contiguous();
}

void check_rc(int rc, const char *what) { TORCH_CHECK(rc == 0, what, " failed (", rc, "): ", pyitd_last_error()); }

// ---- CUDA implementations -----------------------------------------------------------------------------------------
std::tuple<Tensor, Tensor, Tensor> find_knots_cuda(const Tensor &x_, int64_t kinds, int64_t capacity) {
    Tensor x = as_batch(x_);
    c10::cuda::CUDAGuard guard(x.device());
    const int64_t S = x.size(0), N = x.size(1), cap = capacity > 0 ? capacity : N;
    PlanRef plan_ref = get_plan(x.get_device(), S, N, precision_code(x, ""), 0, 2, 0);
    pyitd_plan *plan = plan_ref.get();
    auto iopt = x.options().dtype(at::kInt);
    Tensor knots = at::empty({S, cap}, iopt), count = at::empty({S}, iopt), status = at::empty({S}, iopt);
    check_rc(pyitd_find_knots_device(plan, x.data_ptr(), (int)kinds, knots.data_ptr<int32_t>(), cap, count.data_ptr<int32_t>(),
                                     status.data_ptr<int32_t>(), c10::cuda::getCurrentCUDAStream().stream()),
             "pyitd_find_knots_device");
    return {knots, count, status};
}

std::tuple<Tensor, Tensor, Tensor, Tensor> extract_level_cuda(const Tensor &x_, std::string precision) {
    Tensor x = as_batch(x_);
    c10::cuda::CUDAGuard guard(x.device());
    const int64_t S = x.size(0), N = x.size(1);
    PlanRef plan_ref = get_plan(x.get_device(), S, N, precision_code(x, precision), 0, 2, 0);
    pyitd_plan *plan = plan_ref.get();
    auto iopt = x.options().dtype(at::kInt);
    Tensor R = at::empty_like(x), B = at::empty_like(x), count = at::empty({S}, iopt), status = at::empty({S}, iopt);
    check_rc(pyitd_extract_level_device(plan, x.data_ptr(), R.data_ptr(), B.data_ptr(), count.data_ptr<int32_t>(),
                                        status.data_ptr<int32_t>(), c10::cuda::getCurrentCUDAStream().stream()),
             "pyitd_extract_level_device");
    return {R, B, count, status};
}

// -> (rotations[S, rows, N], n_rows[S], knot_counts[S, rows], baselines[S, rows, N] or an empty tensor, status[S],
//     stop_kind[S], input_knots[S]); rows = max_iteration + 2.  No host synchronisation.
std::tuple<Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor>
decompose_cuda(const Tensor &x_, int64_t max_iteration, int64_t min_extrema, bool return_baselines, bool zero_tail,
               std::string precision) {
    Tensor x = as_batch(x_);
    c10::cuda::CUDAGuard guard(x.device());
    const int64_t S = x.size(0), N = x.size(1);
    const int options = (return_baselines ? PYITD_OPT_BASELINES : 0) | (zero_tail ? PYITD_OPT_ZERO_TAIL : 0);
    PlanRef plan_ref = get_plan(x.get_device(), S, N, precision_code(x, precision), (int)max_iteration, (int)min_extrema, options);
    pyitd_plan *plan = plan_ref.get();
    const int64_t rows = pyitd_plan_rows(plan);
    auto iopt = x.options().dtype(at::kInt);
    Tensor rot = at::empty({S, rows, N}, x.options());
    Tensor bas = return_baselines ? at::empty({S, rows, N}, x.options()) : at::empty({0}, x.options());
    Tensor n_rows = at::empty({S}, iopt), counts = at::empty({S, rows}, iopt), status = at::empty({S}, iopt);
    Tensor kind = at::empty({S}, iopt), iknots = at::empty({S}, iopt);
    check_rc(pyitd_decompose_device(plan, x.data_ptr(), rot.data_ptr(), return_baselines ? bas.data_ptr() : nullptr,
                                    n_rows.data_ptr<int32_t>(), counts.data_ptr<int32_t>(), iknots.data_ptr<int32_t>(),
                                    kind.data_ptr<int32_t>(), status.data_ptr<int32_t>(),
                                    c10::cuda::getCurrentCUDAStream().stream()),
             "pyitd_decompose_device");
    return {rot, n_rows, counts, bas, status, kind, iknots};
}

// ---- Meta implementations (shapes only) ---------------------------------------------------------------------------
std::tuple<Tensor, Tensor, Tensor> find_knots_meta(const Tensor &x, int64_t, int64_t capacity) {
    const int64_t S = x.dim() == 1 ? 1 : x.size(0), N = x.size(-1);
    auto iopt = x.options().dtype(at::kInt);
    return {at::empty({S, capacity > 0 ? capacity : N}, iopt), at::empty({S}, iopt), at::empty({S}, iopt)};
}
std::tuple<Tensor, Tensor, Tensor, Tensor> extract_level_meta(const Tensor &x, std::string) {
    const int64_t S = x.dim() == 1 ? 1 : x.size(0), N = x.size(-1);
    auto iopt = x.options().dtype(at::kInt);
    return {at::empty({S, N}, x.options()), at::empty({S, N}, x.options()), at::empty({S}, iopt), at::empty({S}, iopt)};
}
std::tuple<Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor>
decompose_meta(const Tensor &x, int64_t max_iteration, int64_t, bool return_baselines, bool, std::string) {
    const int64_t S = x.dim() == 1 ? 1 : x.size(0), N = x.size(-1), rows = max_iteration + 2;
    auto iopt = x.options().dtype(at::kInt);
    return {at::empty({S, rows, N}, x.options()), at::empty({S}, iopt), at::empty({S, rows}, iopt),
            return_baselines ? at::empty({S, rows, N}, x.options()) : at::empty({0}, x.options()),
            at::empty({S}, iopt), at::empty({S}, iopt), at::empty({S}, iopt)};
}

int64_t abi_version() { return pyitd_abi_version(); }

}  // namespace

TORCH_LIBRARY(pyitd, m) {
    m.def("find_knots(Tensor x, int kinds=3, int capacity=0) -> (Tensor knots, Tensor count, Tensor status)");
    m.def("extract_level(Tensor x, str precision=\"auto\") -> (Tensor rotation, Tensor baseline, Tensor count, Tensor status)");
    m.def("decompose(Tensor x, int max_iteration=11, int min_extrema=2, bool return_baselines=False, bool zero_tail=False, "
          "str precision=\"auto\") -> (Tensor rotations, Tensor n_rows, Tensor knot_counts, Tensor baselines, Tensor status, "
          "Tensor stop_kind, Tensor input_knots)");
    m.def("clear_plans() -> ()", &clear_plans);
    m.def("abi_version() -> int", &abi_version);
}

TORCH_LIBRARY_IMPL(pyitd, CUDA, m) {
    m.impl("find_knots", &find_knots_cuda);
    m.impl("extract_level", &extract_level_cuda);
    m.impl("decompose", &decompose_cuda);
}

TORCH_LIBRARY_IMPL(pyitd, Meta, m) {
    m.impl("find_knots", &find_knots_meta);
    m.impl("extract_level", &extract_level_meta);
    m.impl("decompose", &decompose_meta);
}
