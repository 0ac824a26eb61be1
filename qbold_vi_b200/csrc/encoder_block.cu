// Streaming kernels of the encoder's fused training blocks -- launchers and C entry points (kernels:
// encoder_block_kernels.cuh).
#include "encoder_block_kernels.cuh"

using namespace qb;

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int64_t stream_grid4(int64_t total4) {
    const int64_t want = (total4 + kThreads - 1) / kThreads, cap = (int64_t)sm_count() * 16;
    return want < cap ? (want < 1 ? 1 : want) : cap;
}

extern "C" int qbold_block_mix_forward(const float* skip, const float* r0, const float* r_bias, const float* z,
                                       float offset, int64_t n, int32_t channels, float* out, float* out_relu,
                                       void* stream) {
    if (n < 0 || channels < 4 || (channels & 3) || channels > 4 * kMaxC4)
        return fail(QBOLD_EUNSUPPORTED, "qbold_block_mix_forward: channels must be a multiple of 4 in [4, 64]");
    if (n == 0) return QBOLD_OK;
    if (!skip || !r0 || !z || !out) return fail(QBOLD_EINVAL, "qbold_block_mix_forward: null pointer");
    if (!aligned16(skip) || !aligned16(r0) || !aligned16(z) || !aligned16(out) || (r_bias && !aligned16(r_bias)) ||
        (out_relu && !aligned16(out_relu)))
        return fail(QBOLD_EINVAL, "qbold_block_mix_forward: operands must be 16-byte aligned");
    const int64_t total4 = n * (channels / 4);
    k_block_mix_fwd<<<(unsigned)stream_grid4(total4), kThreads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(skip), reinterpret_cast<const float4*>(r0),
        reinterpret_cast<const float4*>(r_bias), reinterpret_cast<const float4*>(z), offset, total4, channels / 4,
        reinterpret_cast<float4*>(out), reinterpret_cast<float4*>(out_relu));
    return after_launch("k_block_mix_fwd");
}

extern "C" int qbold_block_mix_backward_add(const float* go, const float* skip, const float* r0, const float* r_bias,
                                            const float* z, float offset, int64_t n, int32_t channels,
                                            int32_t skip_is_relu, const float* skip_addend, float* d_skip, float* d_r,
                                            float* d_z, void* stream);

extern "C" int qbold_block_mix_backward(const float* go, const float* skip, const float* r0, const float* r_bias,
                                        const float* z, float offset, int64_t n, int32_t channels,
                                        int32_t skip_is_relu, float* d_skip, float* d_r, float* d_z, void* stream) {
    return qbold_block_mix_backward_add(go, skip, r0, r_bias, z, offset, n, channels, skip_is_relu, nullptr, d_skip, d_r,
                                        d_z, stream);
}

extern "C" int qbold_block_mix_backward_add(const float* go, const float* skip, const float* r0, const float* r_bias,
                                            const float* z, float offset, int64_t n, int32_t channels,
                                            int32_t skip_is_relu, const float* skip_addend, float* d_skip, float* d_r,
                                            float* d_z, void* stream) {
    if (n < 0 || channels < 4 || (channels & 3) || channels > 4 * kMaxC4)
        return fail(QBOLD_EUNSUPPORTED, "qbold_block_mix_backward: channels must be a multiple of 4 in [4, 64]");
    if (n == 0) return QBOLD_OK;
    if (!go || !skip || !r0 || !z || !d_skip || !d_r || !d_z)
        return fail(QBOLD_EINVAL, "qbold_block_mix_backward: null pointer");
    const void* ptrs[] = {go, skip, r0, z, d_skip, d_r, d_z};
    for (const void* p : ptrs)
        if (!aligned16(p)) return fail(QBOLD_EINVAL, "qbold_block_mix_backward: operands must be 16-byte aligned");
    if (r_bias && !aligned16(r_bias)) return fail(QBOLD_EINVAL, "qbold_block_mix_backward: r_bias must be 16-byte aligned");
    if (skip_addend && !aligned16(skip_addend))
        return fail(QBOLD_EINVAL, "qbold_block_mix_backward: skip_addend must be 16-byte aligned");
    const int64_t total4 = n * (channels / 4);
    k_block_mix_bwd<<<(unsigned)stream_grid4(total4), kThreads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(go), reinterpret_cast<const float4*>(skip), reinterpret_cast<const float4*>(r0),
        reinterpret_cast<const float4*>(r_bias), reinterpret_cast<const float4*>(z), offset, total4, channels / 4,
        skip_is_relu, reinterpret_cast<const float4*>(skip_addend), reinterpret_cast<float4*>(d_skip),
        reinterpret_cast<float4*>(d_r), reinterpret_cast<float4*>(d_z));
    return after_launch("k_block_mix_bwd");
}

extern "C" int64_t qbold_colsum_workspace_floats(void) { return (int64_t)sm_count() * 8 * 4 * kMaxC4; }

// out = g * [y > 0] (+ addend) (y NULL: no mask, out unused) and colsum[channels] (+)= its column sums (may be NULL).
extern "C" int qbold_relu_bwd_colsum(const float* g, const float* y, const float* addend, int64_t n, int32_t channels,
                                     float* out, float* colsum, int32_t accumulate, float* workspace, void* stream) {
    if (n < 0 || channels < 4 || (channels & 3) || channels > 4 * kMaxC4)
        return fail(QBOLD_EUNSUPPORTED, "qbold_relu_bwd_colsum: channels must be a multiple of 4 in [4, 64]");
    if (!g || (y && !out && !colsum) || (!y && !colsum) || (colsum && !workspace))
        return fail(QBOLD_EINVAL, "qbold_relu_bwd_colsum: null pointer");
    if (addend && !y) return fail(QBOLD_EINVAL, "qbold_relu_bwd_colsum: addend needs the masked form (y != NULL)");
    if (!aligned16(g) || (y && !aligned16(y)) || (out && !aligned16(out)) || (workspace && !aligned16(workspace)) ||
        (addend && !aligned16(addend)))
        return fail(QBOLD_EINVAL, "qbold_relu_bwd_colsum: operands must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int c4 = channels / 4;
    int64_t grid = (int64_t)sm_count() * 8;
    const int64_t rows_per_pass = kColTile / c4;
    const int64_t want = (n + rows_per_pass * 4 - 1) / (rows_per_pass * 4);
    if (want < grid) grid = want < 1 ? 1 : want;
    if (n > 0) {
        if (y)
            k_relu_bwd_colsum<true><<<(unsigned)grid, kColTile, 0, st>>>(
                reinterpret_cast<const float4*>(g), reinterpret_cast<const float4*>(y),
                reinterpret_cast<const float4*>(addend), n, c4, reinterpret_cast<float4*>(out),
                colsum ? workspace : nullptr);
        else
            k_relu_bwd_colsum<false><<<(unsigned)grid, kColTile, 0, st>>>(reinterpret_cast<const float4*>(g), nullptr, nullptr,
                                                                          n, c4, nullptr, workspace);
        int rc = after_launch("k_relu_bwd_colsum");
        if (rc) return rc;
    }
    if (colsum) {
        k_colsum_finish<<<1, 1024, 0, st>>>(workspace, n > 0 ? (int)grid : 0, channels, colsum, accumulate);
        return after_launch("k_colsum_finish");
    }
    return QBOLD_OK;
}

extern "C" int qbold_normalise_zouter(const float* data, int64_t b, int32_t nx, int32_t ny, int32_t nz, int32_t n_tau,
                                      int32_t se_idx, int32_t multi_image_normalisation, float* out, void* stream) {
    if (b < 0 || nx < 1 || ny < 1 || nz < 1 || n_tau < 1 || n_tau > 64 || se_idx < 0 || se_idx >= n_tau ||
        (multi_image_normalisation && (se_idx < 1 || se_idx + 1 >= n_tau)))
        return fail(QBOLD_EINVAL, "qbold_normalise_zouter: bad shape or se_idx");
    if (b == 0) return QBOLD_OK;
    if (!data || !out) return fail(QBOLD_EINVAL, "qbold_normalise_zouter: null pointer");
    const int64_t blocks = b * nx * ((ny + 31) / 32) * ((nz + 31) / 32);
    if (blocks > 0x7fffffffLL) return fail(QBOLD_EUNSUPPORTED, "qbold_normalise_zouter: volume too large for one launch");
    const int tp = (n_tau + 3) & ~3;
    const size_t smem = (size_t)32 * (32 * n_tau + 1) * sizeof(float);
    if (smem > 48 * 1024) {
        int rc = cuda_check(cudaFuncSetAttribute(k_normalise_zouter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                            "cudaFuncSetAttribute(k_normalise_zouter)");
        if (rc) return rc;
    }
    k_normalise_zouter<<<(unsigned)blocks, 256, smem, (cudaStream_t)stream>>>(data, nx, ny, nz, n_tau, tp, se_idx,
                                                                              multi_image_normalisation, out);
    return after_launch("k_normalise_zouter");
}
