// K1 / K1b: forward signal model and forward + vector-Jacobian product.
// Replaces SignalGenerationLayer.call (reference signals.py:55-114) and what
// tape.gradient does through it (bessel_j0' = -bessel_j1, node 0 value-dead / gradient-live).
// The kernels themselves live in forward_kernels.cuh; this file holds their launchers, the log-linear kernel and the
// C entry points.
#include "forward_kernels.cuh"

namespace qb {

template <bool BWD>
static int launch_forward_pair(const QboldParams* p, const float* oef_dbv, const float* g, float* signal, float* grad,
                               int64_t n, cudaStream_t st) {
    static int blocks_per_sm = 0;
    if (blocks_per_sm == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_forward_pair<BWD>, kThreads, 0) !=
                cudaSuccess || blocks_per_sm < 1)
            blocks_per_sm = 1;
    }
    const int64_t want = ((n + 1) / 2 + (kThreads / 32) - 1) / (kThreads / 32);
    int64_t grid = (int64_t)sm_count() * blocks_per_sm;
    if (want < grid) grid = want;
    if (grid < 1) grid = 1;
    unsigned long long* work = next_work_counter(st);
    if (!work) return fail(QBOLD_ECUDA, "qbold_forward: work counter unavailable");
    k_forward_pair<BWD><<<(unsigned)grid, kThreads, 0, st>>>(*p, oef_dbv, g, signal, grad, n, work);
    return after_launch("k_forward_pair");
}

template <bool BWD>
static int launch_loglinear(const QboldParams* p, const float* oef_dbv, const float* g, float* signal, float* grad,
                            int64_t n, cudaStream_t st) {
    const size_t smem = sizeof(float) * kThreads * p->n_tau;
    const int64_t grid = (n + kThreads - 1) / kThreads;
    k_loglinear<BWD><<<(unsigned)grid, kThreads, smem, st>>>(*p, reinterpret_cast<const float2*>(oef_dbv), g, signal,
                                                             reinterpret_cast<float2*>(grad), n);
    return after_launch("k_loglinear");
}

template <bool BWD, bool HCT, int PATH>
static int launch_forward_t(const QboldParams* p, const float* oef_dbv, const float* g, float* signal,
                          float* grad, int64_t n, cudaStream_t st) {
    static int blocks_per_sm = 0;
    if (blocks_per_sm == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_forward<BWD, HCT, PATH>, kThreads, 0) !=
                cudaSuccess || blocks_per_sm < 1)
            blocks_per_sm = 1;
    }
    const int64_t want = (n + (kThreads / 32) - 1) / (kThreads / 32);
    int64_t grid = (int64_t)sm_count() * blocks_per_sm;
    if (want < grid) grid = want;
    if (grid < 1) grid = 1;
    unsigned long long* work = next_work_counter(st);
    if (!work) return fail(QBOLD_ECUDA, "qbold_forward: work counter unavailable");
    k_forward<BWD, HCT, PATH><<<(unsigned)grid, kThreads, 0, st>>>(*p, oef_dbv, g, signal, grad, n, work);
    return after_launch("k_forward");
}

template <bool BWD, bool HCT>
static int launch_forward(const QboldParams* p, const float* oef_dbv, const float* g, float* signal,
                          float* grad, int64_t n, cudaStream_t st) {
    if (!p->full_model && !HCT) return launch_loglinear<BWD>(p, oef_dbv, g, signal, grad, n, st);
    if (p->sched_phases > 0 && p->full_model && !HCT && p->n_tau <= 16)
        return launch_forward_pair<BWD>(p, oef_dbv, g, signal, grad, n, st);
    if (p->sched_phases > 0) return launch_forward_t<BWD, HCT, kSched>(p, oef_dbv, g, signal, grad, n, st);
    return p->n_cols > kColGroup ? launch_forward_t<BWD, HCT, kColsMulti>(p, oef_dbv, g, signal, grad, n, st)
                                 : launch_forward_t<BWD, HCT, kCols>(p, oef_dbv, g, signal, grad, n, st);
}

template <bool HCT, int PATH>
static int launch_misalign_t(const QboldParams* p, const float* oef_dbv, int64_t n, float prob, const float* u,
                             const int32_t* idx, const float* eps, uint64_t seed, uint64_t offset, float* signal,
                             cudaStream_t st) {
    static int blocks_per_sm = 0;
    if (blocks_per_sm == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_misalign<HCT, PATH>, kThreads, 0) !=
                cudaSuccess || blocks_per_sm < 1)
            blocks_per_sm = 1;
    }
    const int64_t want = (n + (kThreads / 32) - 1) / (kThreads / 32);
    int64_t grid = (int64_t)sm_count() * blocks_per_sm;
    if (want < grid) grid = want;
    if (grid < 1) grid = 1;
    unsigned long long* work = next_work_counter(st);
    if (!work) return fail(QBOLD_ECUDA, "qbold_misalign: work counter unavailable");
    k_misalign<HCT, PATH><<<(unsigned)grid, kThreads, 0, st>>>(*p, oef_dbv, n, prob, u, idx, eps, seed, offset, signal,
                                                              work);
    return after_launch("k_misalign");
}

template <bool HCT>
static int launch_misalign(const QboldParams* p, const float* oef_dbv, int64_t n, float prob, const float* u,
                           const int32_t* idx, const float* eps, uint64_t seed, uint64_t offset, float* signal,
                           cudaStream_t st) {
    if (p->sched_phases > 0) return launch_misalign_t<HCT, kSched>(p, oef_dbv, n, prob, u, idx, eps, seed, offset, signal, st);
    return p->n_cols > kColGroup
               ? launch_misalign_t<HCT, kColsMulti>(p, oef_dbv, n, prob, u, idx, eps, seed, offset, signal, st)
               : launch_misalign_t<HCT, kCols>(p, oef_dbv, n, prob, u, idx, eps, seed, offset, signal, st);
}

}  // namespace qb

using namespace qb;

extern "C" int qbold_misalign(const QboldParams* p, const float* oef_dbv, int32_t width, int64_t n, float prob,
                              const float* sel_u01, const int32_t* from_index, const float* eps, uint64_t seed,
                              uint64_t offset, float* signal, void* stream) {
    if (!p || p->abi_version != QBOLD_ABI_VERSION) return fail(QBOLD_EINVAL, "qbold_misalign: bad params block");
    if (width != 2 && width != 3)
        return fail(QBOLD_EINVAL, "Input should have 2 (OEF, DBV) or 3 (OEF, DBV, hct) elements in last dimension");
    if (n < 0 || (n > 0 && (!oef_dbv || !signal))) return fail(QBOLD_EINVAL, "qbold_misalign: null pointer");
    const bool any = sel_u01 || from_index || eps, all = sel_u01 && from_index && eps;
    if (any && !all)
        return fail(QBOLD_EINVAL, "qbold_misalign: pass all of sel_u01, from_index and eps, or none of them");
    if (p->n_tau - 1 <= 4)
        return fail(QBOLD_EUNSUPPORTED, "misalignment draws the first misaligned image from [4, n_tau - 1): needs "
                                        "n_tau > 5 (signals.py:84-85), got %d", p->n_tau);
    if (n == 0 || !(prob > 0.f)) return QBOLD_OK;
    cudaStream_t st = (cudaStream_t)stream;
    return width == 3 ? launch_misalign<true>(p, oef_dbv, n, prob, sel_u01, from_index, eps, seed, offset, signal, st)
                      : launch_misalign<false>(p, oef_dbv, n, prob, sel_u01, from_index, eps, seed, offset, signal, st);
}

extern "C" int qbold_forward_backward_hct(const QboldParams* p, const float* oef_dbv_hct, const float* g_signal,
                                          int64_t n, float* signal, float* g_oef_dbv_hct, void* stream) {
    if (!p || p->abi_version != QBOLD_ABI_VERSION)
        return fail(QBOLD_EINVAL, "qbold_forward_backward_hct: bad params block");
    if (n < 0 || (n > 0 && (!oef_dbv_hct || !g_oef_dbv_hct)))
        return fail(QBOLD_EINVAL, "qbold_forward_backward_hct: null pointer");
    if (n == 0) return QBOLD_OK;
    return launch_forward<true, true>(p, oef_dbv_hct, g_signal, signal, g_oef_dbv_hct, n, (cudaStream_t)stream);
}

extern "C" int qbold_forward(const QboldParams* p, const float* oef_dbv, int32_t width, int64_t n,
                             float* signal, void* stream) {
    if (!p || p->abi_version != QBOLD_ABI_VERSION) return fail(QBOLD_EINVAL, "qbold_forward: bad params block");
    if (width != 2 && width != 3)
        return fail(QBOLD_EINVAL, "Input should have 2 (OEF, DBV) or 3 (OEF, DBV, hct) elements in last dimension");
    if (n < 0 || (n > 0 && (!oef_dbv || !signal))) return fail(QBOLD_EINVAL, "qbold_forward: null pointer");
    if (n == 0) return QBOLD_OK;
    cudaStream_t st = (cudaStream_t)stream;
    return width == 3 ? launch_forward<false, true>(p, oef_dbv, nullptr, signal, nullptr, n, st)
                      : launch_forward<false, false>(p, oef_dbv, nullptr, signal, nullptr, n, st);
}

extern "C" int qbold_forward_backward(const QboldParams* p, const float* oef_dbv, const float* g_signal,
                                      int64_t n, float* signal, float* g_oef_dbv, void* stream) {
    if (!p || p->abi_version != QBOLD_ABI_VERSION)
        return fail(QBOLD_EINVAL, "qbold_forward_backward: bad params block");
    if (n < 0 || (n > 0 && (!oef_dbv || !g_oef_dbv)))
        return fail(QBOLD_EINVAL, "qbold_forward_backward: null pointer");
    if (n == 0) return QBOLD_OK;
    return launch_forward<true, false>(p, oef_dbv, g_signal, signal, g_oef_dbv, n, (cudaStream_t)stream);
}
