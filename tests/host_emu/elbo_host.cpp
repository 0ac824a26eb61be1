// TEST INFRASTRUCTURE: the likelihood-side kernels (qbold_vi_b200/csrc/elbo_kernels.cuh -- fused ELBO, KL, NLL,
// reparameterised sample, posterior statistics; the source libqbold.so is built from) compiled for the host and run in
// the SIMT emulator of tests/host_emu/shim/cuda_runtime.h.  See forward_host.cpp.
#include <cuda_runtime.h>      // the shim

#include "elbo_kernels.cuh"

namespace {
unsigned long long g_work;
}

// k_elbo_pair<HAS_PRIOR> (the production kernel for n_tau <= 16, full model) or, pair == 0, k_elbo<HAS_PRIOR, PATH>.
// seed_dev non-NULL: the Philox key is read from memory (qbold_elbo_fused_graph); inv_mask_sum_dev likewise.
QB_EMU_API int qb_emu_elbo(const QboldParams* P, const float* q, const float* sigma, const float* y, const float* mask,
                           const float* prior, const float* eps, const float* eps_kl, uint64_t seed,
                           const uint64_t* seed_dev, uint64_t offset, int kl_samples, float inv_mask_sum,
                           const float* inv_mask_sum_dev, float kl_weight, int64_t n, float* grad_q, float* grad_sigma,
                           float* nll_map, float* kl_map, double* sums, int pair, int path, int grid) {
    g_work = 0;
    const QboldParams params = *P;
    const int block = qb::kThreads;
#define QB_ARGS params, q, sigma, y, mask, prior, eps, eps_kl, seed, seed_dev, offset, kl_samples, inv_mask_sum, inv_mask_sum_dev, \
                kl_weight, n, grad_q, grad_sigma, nll_map, kl_map, sums, &g_work
    if (pair) {
        if (prior) qb_emu::launch(grid, block, [&]() { qb::k_elbo_pair<true>(QB_ARGS); });
        else qb_emu::launch(grid, block, [&]() { qb::k_elbo_pair<false>(QB_ARGS); });
        return 0;
    }
#define QB_CASE(HP, PA)                                                              \
    if ((prior != nullptr) == HP && path == PA) {                                    \
        qb_emu::launch(grid, block, [&]() { qb::k_elbo<HP, PA>(QB_ARGS); });         \
        return 0;                                                                    \
    }
    QB_CASE(true, 0) QB_CASE(false, 0) QB_CASE(true, 1) QB_CASE(false, 1) QB_CASE(true, 2) QB_CASE(false, 2)
#undef QB_CASE
#undef QB_ARGS
    return -1;
}

QB_EMU_API void qb_emu_kl(const float* q, const float* prior, const float* mask, const float* eps_kl, uint64_t seed,
                          uint64_t offset, int n_samples, int64_t n, float* kl_map, float* grad_q, int grid) {
    qb_emu::launch(grid, qb::kThreads, [&]() { qb::k_kl(q, prior, mask, eps_kl, seed, offset, n_samples, n, kl_map, grad_q); });
}

QB_EMU_API void qb_emu_reparam(const float* q, const float* eps, uint64_t seed, uint64_t offset, int64_t n, float* out) {
    const int grid = (int)((n + qb::kThreads - 1) / qb::kThreads);
    qb_emu::launch(grid, qb::kThreads, [&]() { qb::k_reparam(q, eps, seed, offset, n, out); });
}

QB_EMU_API void qb_emu_posterior_stats(float dw_k, const float* q, const float* eps, uint64_t seed, uint64_t offset,
                                       int n_samples, int64_t n, float* mean3, float* var3, int grid) {
    qb_emu::launch(grid, qb::kThreads,
                   [&]() { qb::k_posterior_stats(dw_k, q, eps, seed, offset, n_samples, n, mean3, var3); });
}

// likelihood map of save_predictions (model.py:808-817): pair != 0 -> k_nll_map_pair, else k_nll_map<path>
QB_EMU_API int qb_emu_nll_map(const QboldParams* P, const float* q, const float* sigma, const float* y, const float* mask,
                              const float* eps, uint64_t seed, uint64_t offset, int n_samples, int64_t n, float* nll_map,
                              int pair, int path, int grid) {
    g_work = 0;
    const QboldParams params = *P;
    if (pair) {
        qb_emu::launch(grid, qb::kThreads, [&]() {
            qb::k_nll_map_pair(params, q, sigma, y, mask, eps, seed, offset, n_samples, n, nll_map, &g_work);
        });
        return 0;
    }
#define QB_CASE(PA)                                                                                               \
    if (path == PA) {                                                                                             \
        qb_emu::launch(grid, qb::kThreads, [&]() {                                                                \
            qb::k_nll_map<PA>(params, q, sigma, y, mask, eps, seed, offset, n_samples, n, nll_map, &g_work);      \
        });                                                                                                       \
        return 0;                                                                                                 \
    }
    QB_CASE(0) QB_CASE(1) QB_CASE(2)
#undef QB_CASE
    return -1;
}

// fine_tune_loss_fn alone on predictions that already exist (model.py:527-568): W = 16 (two voxels per warp) for
// n_tau <= 16, else 32 -- as qbold_nll picks it
QB_EMU_API void qb_emu_nll(const QboldParams* P, const float* y, const float* pred, const float* sigma, const float* mask,
                           int64_t n, float* nll_map, float* d_pred, float* d_sigma, int grid) {
    const QboldParams params = *P;
    if (params.n_tau <= 16)
        qb_emu::launch(grid, qb::kThreads, [&]() { qb::k_nll<16>(params, y, pred, sigma, mask, n, nll_map, d_pred, d_sigma); });
    else
        qb_emu::launch(grid, qb::kThreads, [&]() { qb::k_nll<32>(params, y, pred, sigma, mask, n, nll_map, d_pred, d_sigma); });
}
