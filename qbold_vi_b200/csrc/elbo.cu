// K2: fused amortized-VI training step of the likelihood side -- launchers and C entry points.
// The kernels (and the description of what they replace in the reference) are in elbo_kernels.cuh.
#include "elbo_kernels.cuh"

namespace qb {

template <typename K>
static int64_t persistent_grid(K kernel, int64_t n_warp_items) {
    int bps = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kernel, kThreads, 0) != cudaSuccess || bps < 1) bps = 1;
    const int64_t want = (n_warp_items + (kThreads / 32) - 1) / (kThreads / 32);
    int64_t grid = (int64_t)sm_count() * bps;
    if (want < grid) grid = want;
    return grid < 1 ? 1 : grid;
}

}  // namespace qb

using namespace qb;

static int elbo_fused_impl(const QboldParams* p, const float* q, const float* sigma, const float* y,
                           const float* mask, const float* prior, const float* eps, const float* eps_kl,
                           uint64_t seed, const uint64_t* seed_dev, uint64_t offset, int32_t kl_samples,
                           float inv_mask_sum, const float* inv_mask_sum_dev, float kl_weight, int64_t n,
                           float* grad_q, float* grad_sigma, float* nll_map, float* kl_map, double* sums,
                           void* stream) {
    if (!p || p->abi_version != QBOLD_ABI_VERSION) return fail(QBOLD_EINVAL, "qbold_elbo_fused: bad params block");
    if (n < 0 || kl_samples < 0) return fail(QBOLD_EINVAL, "qbold_elbo_fused: negative size");
    if (n == 0) return QBOLD_OK;
    if (!q || !sigma || !y || !mask || !grad_q || !grad_sigma)
        return fail(QBOLD_EINVAL, "qbold_elbo_fused: null pointer");
    if (p->n_tau > 32) return fail(QBOLD_EINVAL, "qbold_elbo_fused: n_tau > 32");
    cudaStream_t st = (cudaStream_t)stream;
    const int path = p->sched_phases > 0 ? kSched : (p->n_cols > kColGroup ? kColsMulti : kCols);
    const int64_t want = (n + 7) / 8;
    unsigned long long* work = next_work_counter(st);
    if (!work) return fail(QBOLD_ECUDA, "qbold_elbo_fused: work counter unavailable");
#define QB_LAUNCH_ELBO(HP, MU)                                                                                       \
    do {                                                                                                              \
        static int64_t grid_cache = 0;                                                                                \
        const int64_t grid = grid_cache ? grid_cache : (grid_cache = persistent_grid(k_elbo<HP, MU>, INT64_MAX / 64)); \
        k_elbo<HP, MU><<<(unsigned)(want < grid ? want : grid), kThreads, 0, st>>>(                                  \
            *p, q, sigma, y, mask, HP ? prior : nullptr, eps, HP ? eps_kl : nullptr, seed, seed_dev, offset,          \
            HP ? kl_samples : 0, inv_mask_sum, inv_mask_sum_dev, kl_weight, n, grad_q, grad_sigma, nll_map, kl_map, sums, work);        \
    } while (0)
    if (path == kSched && p->full_model && p->n_tau <= 16) {
        const int64_t wantp = ((n + 31) / 32 + 7) / 8;                      // a warp takes units of 32 voxels
        if (prior) {
            static int64_t grid_cache = 0;
            const int64_t grid = grid_cache ? grid_cache : (grid_cache = persistent_grid(k_elbo_pair<true>, INT64_MAX / 64));
            k_elbo_pair<true><<<(unsigned)(wantp < grid ? wantp : grid), kThreads, 0, st>>>(
                *p, q, sigma, y, mask, prior, eps, eps_kl, seed, seed_dev, offset, kl_samples, inv_mask_sum, inv_mask_sum_dev,
                kl_weight, n, grad_q,
                grad_sigma, nll_map, kl_map, sums, work);
        } else {
            static int64_t grid_cache = 0;
            const int64_t grid = grid_cache ? grid_cache : (grid_cache = persistent_grid(k_elbo_pair<false>, INT64_MAX / 64));
            k_elbo_pair<false><<<(unsigned)(wantp < grid ? wantp : grid), kThreads, 0, st>>>(
                *p, q, sigma, y, mask, nullptr, eps, nullptr, seed, seed_dev, offset, 0, inv_mask_sum, inv_mask_sum_dev, kl_weight,
                n, grad_q,
                grad_sigma, nll_map, kl_map, sums, work);
        }
    } else if (prior) {
        if (path == kSched) QB_LAUNCH_ELBO(true, kSched);
        else if (path == kCols) QB_LAUNCH_ELBO(true, kCols);
        else QB_LAUNCH_ELBO(true, kColsMulti);
    } else {
        if (path == kSched) QB_LAUNCH_ELBO(false, kSched);
        else if (path == kCols) QB_LAUNCH_ELBO(false, kCols);
        else QB_LAUNCH_ELBO(false, kColsMulti);
    }
#undef QB_LAUNCH_ELBO
    return after_launch("k_elbo");
}

extern "C" int qbold_elbo_fused(const QboldParams* p, const float* q, const float* sigma, const float* y,
                                const float* mask, const float* prior, const float* eps, const float* eps_kl,
                                uint64_t seed, uint64_t offset, int32_t kl_samples, float inv_mask_sum,
                                float kl_weight, int64_t n, float* grad_q, float* grad_sigma, float* nll_map,
                                float* kl_map, double* sums, void* stream) {
    return elbo_fused_impl(p, q, sigma, y, mask, prior, eps, eps_kl, seed, nullptr, offset, kl_samples, inv_mask_sum, nullptr,
                           kl_weight, n, grad_q, grad_sigma, nll_map, kl_map, sums, stream);
}

extern "C" int qbold_elbo_fused_dev(const QboldParams* p, const float* q, const float* sigma, const float* y,
                                    const float* mask, const float* prior, const float* eps, const float* eps_kl,
                                    uint64_t seed, uint64_t offset, int32_t kl_samples,
                                    const float* inv_mask_sum_dev, float kl_weight, int64_t n, float* grad_q,
                                    float* grad_sigma, float* nll_map, float* kl_map, double* sums, void* stream) {
    if (!inv_mask_sum_dev) return fail(QBOLD_EINVAL, "qbold_elbo_fused_dev: inv_mask_sum_dev is NULL");
    return elbo_fused_impl(p, q, sigma, y, mask, prior, eps, eps_kl, seed, nullptr, offset, kl_samples, 0.f, inv_mask_sum_dev,
                           kl_weight, n, grad_q, grad_sigma, nll_map, kl_map, sums, stream);
}

extern "C" int qbold_elbo_fused_graph(const QboldParams* p, const float* q, const float* sigma, const float* y,
                                      const float* mask, const float* prior, const uint64_t* seed_dev, uint64_t offset,
                                      int32_t kl_samples, const float* inv_mask_sum_dev, float kl_weight, int64_t n,
                                      float* grad_q, float* grad_sigma, float* nll_map, float* kl_map, double* sums,
                                      void* stream) {
    if (!inv_mask_sum_dev || !seed_dev)
        return fail(QBOLD_EINVAL, "qbold_elbo_fused_graph: seed_dev / inv_mask_sum_dev is NULL");
    return elbo_fused_impl(p, q, sigma, y, mask, prior, nullptr, nullptr, 0, seed_dev, offset, kl_samples, 0.f,
                           inv_mask_sum_dev, kl_weight, n, grad_q, grad_sigma, nll_map, kl_map, sums, stream);
}

extern "C" int qbold_kl(const float* q, const float* prior, const float* mask, const float* eps_kl, uint64_t seed,
                        uint64_t offset, int32_t n_samples, int64_t n, float* kl_map, float* grad_q, void* stream) {
    if (n < 0 || n_samples < 0 || (n > 0 && (!q || !prior || !kl_map)))
        return fail(QBOLD_EINVAL, "qbold_kl: bad argument");
    if (n == 0) return QBOLD_OK;
    // one thread per voxel; 64-thread granularity spreads a small (masked) batch over all SMs
    const int64_t blocks = (n + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)sm_count() * 64;
    k_kl<<<(unsigned)(blocks < cap ? blocks : cap), kThreads, 0, (cudaStream_t)stream>>>(q, prior, mask, eps_kl, seed,
                                                                                        offset, n_samples, n, kl_map,
                                                                                        grad_q);
    return after_launch("k_kl");
}

extern "C" int qbold_nll_map(const QboldParams* p, const float* q, const float* sigma, const float* y, const float* mask,
                             const float* eps, uint64_t seed, uint64_t offset, int32_t n_samples, int64_t n,
                             float* nll_map, void* stream) {
    if (!p || p->abi_version != QBOLD_ABI_VERSION) return fail(QBOLD_EINVAL, "qbold_nll_map: bad params block");
    if (n < 0 || n_samples < 1) return fail(QBOLD_EINVAL, "qbold_nll_map: bad size");
    if (n == 0) return QBOLD_OK;
    if (!q || !sigma || !y || !nll_map) return fail(QBOLD_EINVAL, "qbold_nll_map: null pointer");
    const int path = p->sched_phases > 0 ? kSched : (p->n_cols > kColGroup ? kColsMulti : kCols);
    const int64_t want = (n + 7) / 8;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* work = next_work_counter(st);
    if (!work) return fail(QBOLD_ECUDA, "qbold_nll_map: work counter unavailable");
#define QB_LAUNCH_NLL(PA)                                                                                            \
    do {                                                                                                              \
        static int64_t grid_cache = 0;                                                                                \
        const int64_t grid = grid_cache ? grid_cache : (grid_cache = persistent_grid(k_nll_map<PA>, INT64_MAX / 64)); \
        k_nll_map<PA><<<(unsigned)(want < grid ? want : grid), kThreads, 0, st>>>(*p, q, sigma, y, mask, eps, seed,  \
                                                                                 offset, n_samples, n, nll_map, work); \
    } while (0)
    if (path == kSched && p->full_model && p->n_tau <= 16) {
        static int64_t grid_cache = 0;
        const int64_t grid = grid_cache ? grid_cache : (grid_cache = persistent_grid(k_nll_map_pair, INT64_MAX / 64));
        k_nll_map_pair<<<(unsigned)(want < grid ? want : grid), kThreads, 0, st>>>(*p, q, sigma, y, mask, eps, seed,
                                                                                  offset, n_samples, n, nll_map, work);
    } else if (path == kSched) QB_LAUNCH_NLL(kSched);
    else if (path == kCols) QB_LAUNCH_NLL(kCols);
    else QB_LAUNCH_NLL(kColsMulti);
#undef QB_LAUNCH_NLL
    return after_launch("k_nll_map");
}

extern "C" int qbold_nll(const QboldParams* p, const float* y, const float* pred, const float* sigma, const float* mask,
                         int64_t n, float* nll_map, float* d_pred, float* d_sigma, void* stream) {
    if (!p || p->abi_version != QBOLD_ABI_VERSION) return fail(QBOLD_EINVAL, "qbold_nll: bad params block");
    if (n < 0 || (n > 0 && (!y || !pred || !sigma || !nll_map))) return fail(QBOLD_EINVAL, "qbold_nll: bad argument");
    if (n == 0) return QBOLD_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (p->n_tau <= 16) {
        const int64_t want = ((n + 1) / 2 + 7) / 8;
        k_nll<16><<<(unsigned)(want < cap ? want : cap), kThreads, 0, st>>>(*p, y, pred, sigma, mask, n, nll_map, d_pred,
                                                                           d_sigma);
    } else {
        const int64_t want = (n + 7) / 8;
        k_nll<32><<<(unsigned)(want < cap ? want : cap), kThreads, 0, st>>>(*p, y, pred, sigma, mask, n, nll_map, d_pred,
                                                                           d_sigma);
    }
    return after_launch("k_nll");
}

extern "C" int qbold_reparam_sample(const float* q, const float* eps, uint64_t seed, uint64_t offset, int64_t n,
                                    float* oef_dbv, void* stream) {
    if (n < 0 || (n > 0 && (!q || !oef_dbv))) return fail(QBOLD_EINVAL, "qbold_reparam_sample: bad argument");
    if (n == 0) return QBOLD_OK;
    k_reparam<<<(unsigned)((n + kThreads - 1) / kThreads), kThreads, 0, (cudaStream_t)stream>>>(q, eps, seed, offset,
                                                                                                n, oef_dbv);
    return after_launch("k_reparam");
}

extern "C" int qbold_posterior_stats(const QboldParams* p, const float* q, const float* eps, uint64_t seed,
                                     uint64_t offset, int32_t n_samples, int64_t n, float* mean3, float* var3,
                                     void* stream) {
    if (!p || p->abi_version != QBOLD_ABI_VERSION) return fail(QBOLD_EINVAL, "qbold_posterior_stats: bad params block");
    if (n_samples < 1 || n_samples > 256) return fail(QBOLD_EINVAL, "qbold_posterior_stats: n_samples must be in [1,256]");
    if (n < 0 || (n > 0 && (!q || !mean3 || !var3))) return fail(QBOLD_EINVAL, "qbold_posterior_stats: null pointer");
    if (n == 0) return QBOLD_OK;
    const int64_t blocks = (n + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)sm_count() * 64;
    k_posterior_stats<<<(unsigned)(blocks < cap ? blocks : cap), kThreads, 0, (cudaStream_t)stream>>>(
        p->dw_k, q, eps, seed, offset, n_samples, n, mean3, var3);
    return after_launch("k_posterior_stats");
}
