// TEST INFRASTRUCTURE: the loss kernels either side of the fused path (qbold_vi_b200/csrc/losses_kernels.cuh: TV stencil,
// pre-training NLL, diagonal and mixture-of-Gaussians KL; the source libqbold.so is built from) compiled for the host
// and run in the SIMT emulator.  The option block of k_synth_nll is filled as qbold_synth_nll (losses.cu) fills it.
#include <cuda_runtime.h>      // the shim

#include <cmath>

#include "losses_kernels.cuh"

QB_EMU_API void qb_emu_smoothness(const float* q, int n_ch, const float* mask, int64_t n_vol, int X, int Y, int Z,
                                  float scale, const float* scale_dev, double* tv_sum, float* grad_q, int grid) {
    const int64_t n = n_vol * X * Y * Z;
    qb_emu::launch(grid, qb::kThreads,
                   [&]() { qb::k_smoothness(q, n_ch, mask, n, X, Y, Z, scale, scale_dev, tv_sum, grad_q); });
}

QB_EMU_API void qb_emu_synth_nll(const float* labels, int label_stride, const float* pred, int pred_stride, int use_mvg,
                                 double ig_alpha, double ig_beta, const float* ig4, int64_t n, float grad_scale,
                                 float* nll_rows, float* grad_pred, double* loss_sum, double* ig_sums, int grid) {
    qb::SynthOpts opt{};
    opt.use_mvg = use_mvg ? 1 : 0;
    opt.pred_stride = pred_stride;
    if (ig4 != nullptr) {
        opt.inv_gamma = 1;
    } else if (ig_alpha * ig_beta > 0.0) {
        opt.inv_gamma = 1;
        opt.ig_alpha = (float)ig_alpha;
        opt.ig_beta = (float)ig_beta;
        opt.ig_const = (float)(ig_alpha * std::log(ig_beta) - std::lgamma(ig_alpha));
    }
    qb_emu::launch(grid, qb::kThreads, [&]() {
        qb::k_synth_nll(labels, label_stride, pred, opt, n, grad_scale, nll_rows, grad_pred, loss_sum, ig4, ig_sums);
    });
}

QB_EMU_API void qb_emu_diag_kl(const float* pred, int pred_stride, const float* prior, int prior_stride, const float* mask,
                               int64_t n, float* kl_map, float* grad_pred, int gpred_stride, float* grad_prior,
                               int gprior_stride, int grid) {
    qb_emu::launch(grid, qb::kThreads, [&]() {
        qb::k_diag_kl(pred, pred_stride, prior, prior_stride, mask, n, kl_map, grad_pred, gpred_stride, grad_prior,
                      gprior_stride);
    });
}

QB_EMU_API void qb_emu_mog_kl(const float* pred, int n_comp, const float* mask, const float* eps, uint64_t seed,
                              uint64_t offset, int64_t n, float* kl_map, float* grad_pred, int grid) {
    qb_emu::launch(grid, qb::kThreads,
                   [&]() { qb::k_mog_kl(pred, n_comp, mask, eps, seed, offset, n, kl_map, grad_pred); });
}
