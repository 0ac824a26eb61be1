// TEST INFRASTRUCTURE: the streaming kernels of the encoder's fused training blocks
// (qbold_vi_b200/csrc/encoder_block_kernels.cuh; the source libqbold.so is built from) compiled for the host and run in
// the SIMT emulator.  The functions restate the launch arithmetic of the entry points in encoder_block.cu.
#include <cuda_runtime.h>      // the shim

#include <vector>

#include "encoder_block_kernels.cuh"

using namespace qb;

QB_EMU_API void qb_emu_block_mix_forward(const float* skip, const float* r0, const float* r_bias, const float* z, float offset,
                                         int64_t n, int channels, float* out, float* out_relu, int grid) {
    const int64_t total4 = n * (channels / 4);
    qb_emu::launch(grid, kThreads, [&]() {
        k_block_mix_fwd(reinterpret_cast<const float4*>(skip), reinterpret_cast<const float4*>(r0),
                        reinterpret_cast<const float4*>(r_bias), reinterpret_cast<const float4*>(z), offset, total4,
                        channels / 4, reinterpret_cast<float4*>(out), reinterpret_cast<float4*>(out_relu));
    });
}

QB_EMU_API void qb_emu_block_mix_backward(const float* go, const float* skip, const float* r0, const float* r_bias,
                                          const float* z, float offset, int64_t n, int channels, int skip_is_relu,
                                          const float* skip_addend, float* d_skip, float* d_r, float* d_z, int grid) {
    const int64_t total4 = n * (channels / 4);
    qb_emu::launch(grid, kThreads, [&]() {
        k_block_mix_bwd(reinterpret_cast<const float4*>(go), reinterpret_cast<const float4*>(skip),
                        reinterpret_cast<const float4*>(r0), reinterpret_cast<const float4*>(r_bias),
                        reinterpret_cast<const float4*>(z), offset, total4, channels / 4, skip_is_relu,
                        reinterpret_cast<const float4*>(skip_addend), reinterpret_cast<float4*>(d_skip),
                        reinterpret_cast<float4*>(d_r), reinterpret_cast<float4*>(d_z));
    });
}

// out = g * [y > 0] (+ addend) and colsum (+)= its column sums: two-stage fixed-order reduction (qbold_relu_bwd_colsum)
QB_EMU_API void qb_emu_relu_bwd_colsum(const float* g, const float* y, const float* addend, int64_t n, int channels,
                                       float* out, float* colsum, int accumulate, int grid) {
    const int c4 = channels / 4;
    std::vector<float> ws((size_t)grid * 4 * kMaxC4 + 4, 0.f);
    float* workspace = ws.data();
    if (y)
        qb_emu::launch(grid, kColTile, [&]() {
            k_relu_bwd_colsum<true>(reinterpret_cast<const float4*>(g), reinterpret_cast<const float4*>(y),
                                    reinterpret_cast<const float4*>(addend), n, c4, reinterpret_cast<float4*>(out),
                                    colsum ? workspace : nullptr);
        });
    else
        qb_emu::launch(grid, kColTile, [&]() {
            k_relu_bwd_colsum<false>(reinterpret_cast<const float4*>(g), nullptr, nullptr, n, c4, nullptr, workspace);
        });
    if (colsum) qb_emu::launch(1, 1024, [&]() { k_colsum_finish(workspace, grid, channels, colsum, accumulate); });
}

QB_EMU_API void qb_emu_normalise_zouter(const float* data, int64_t b, int nx, int ny, int nz, int n_tau, int se_idx, int multi,
                                        float* out) {
    const int64_t blocks = b * nx * ((ny + 31) / 32) * ((nz + 31) / 32);
    const int tp = (n_tau + 3) & ~3;
    const size_t smem = (size_t)32 * (32 * n_tau + 1) * sizeof(float);
    qb_emu::launch((int)blocks, 256, [&]() { k_normalise_zouter(data, nx, ny, nz, n_tau, tp, se_idx, multi, out); }, 1, smem);
}

// ---- skinny Dense layers of the heads (encoder_small_kernels.cuh), launch arithmetic of qbold_dense_small_*
#include "encoder_small_kernels.cuh"

QB_EMU_API void qb_emu_dense_small_forward(const float* x, const float* w, const float* bias, int n_in, int n_out, int64_t n,
                                           float* y, int grid) {
    qb_emu::launch(grid, 128, [&]() { k_dense_small_fwd_coop(x, w, bias, n_in, n_out, n, y); });
}

QB_EMU_API void qb_emu_dense_small_dgrad(const float* g, const float* w, const float* relu_mask, int n_in, int n_out,
                                         int64_t n, float* dx, int grid) {
    qb_emu::launch(grid, 256, [&]() { k_dense_small_dgrad_coop(g, w, relu_mask, n_in, n_out, n, dx); });
}
