// Losses either side of the fused path -- launchers and C entry points (kernels: losses_kernels.cuh), and the gated
// residual mix kernels of the encoder blocks.
#include "losses_kernels.cuh"

namespace qb {
namespace {

int64_t stream_grid(int64_t n) {
    const int64_t want = (n + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)sm_count() * 16;          // 16 x 256 threads = 2 full waves of resident CTAs per SM
    return want < cap ? want : cap;
}

}  // namespace
}  // namespace qb

using namespace qb;

static int smoothness_impl(const float* q, int32_t n_ch, const float* mask, int64_t n_vol, int32_t nx, int32_t ny,
                           int32_t nz, float scale, const float* scale_dev, double* tv_sum, float* grad_q,
                           void* stream) {
    if ((n_ch != 4 && n_ch != 5) || n_vol < 0 || nx <= 0 || ny <= 0 || nz <= 0)
        return fail(QBOLD_EINVAL, "qbold_smoothness: bad shape (n_ch=%d, volumes=%lld, %d x %d x %d)", n_ch,
                    (long long)n_vol, nx, ny, nz);
    const int64_t n = n_vol * nx * ny * nz;
    if (n == 0) return QBOLD_OK;
    if (!q || !mask || (!tv_sum && !grad_q)) return fail(QBOLD_EINVAL, "qbold_smoothness: null pointer");
    k_smoothness<<<(unsigned)stream_grid(n), kThreads, 0, (cudaStream_t)stream>>>(q, n_ch, mask, n, nx, ny, nz, scale,
                                                                                  scale_dev, tv_sum, grad_q);
    return after_launch("k_smoothness");
}

extern "C" int qbold_smoothness(const float* q, int32_t n_ch, const float* mask, int64_t n_vol, int32_t nx,
                                int32_t ny, int32_t nz, float scale, double* tv_sum, float* grad_q, void* stream) {
    return smoothness_impl(q, n_ch, mask, n_vol, nx, ny, nz, scale, nullptr, tv_sum, grad_q, stream);
}

extern "C" int qbold_smoothness_dev(const float* q, int32_t n_ch, const float* mask, int64_t n_vol, int32_t nx,
                                    int32_t ny, int32_t nz, const float* scale_dev, double* tv_sum, float* grad_q,
                                    void* stream) {
    if (!scale_dev) return fail(QBOLD_EINVAL, "qbold_smoothness_dev: scale_dev is NULL");
    return smoothness_impl(q, n_ch, mask, n_vol, nx, ny, nz, 0.f, scale_dev, tv_sum, grad_q, stream);
}

extern "C" int qbold_synth_nll(const float* labels, int32_t label_stride, const float* pred, int32_t use_mvg,
                               double inv_gamma_alpha, double inv_gamma_beta, int64_t n, float grad_scale,
                               float* nll_rows, float* grad_pred, double* loss_sum, void* stream) {
    if (n < 0 || label_stride < 2) return fail(QBOLD_EINVAL, "qbold_synth_nll: bad argument");
    if (n == 0) return QBOLD_OK;
    if (!labels || !pred || (!nll_rows && !grad_pred && !loss_sum))
        return fail(QBOLD_EINVAL, "qbold_synth_nll: null pointer");
    SynthOpts opt{};
    opt.use_mvg = use_mvg ? 1 : 0;
    opt.pred_stride = use_mvg ? 5 : 4;
    if (inv_gamma_alpha * inv_gamma_beta > 0.0) {
        opt.inv_gamma = 1;
        opt.ig_alpha = (float)inv_gamma_alpha;
        opt.ig_beta = (float)inv_gamma_beta;
        opt.ig_const = (float)(inv_gamma_alpha * log(inv_gamma_beta) - lgamma(inv_gamma_alpha));
    }
    k_synth_nll<<<(unsigned)stream_grid(n), kThreads, 0, (cudaStream_t)stream>>>(labels, label_stride, pred, opt, n,
                                                                                 grad_scale, nll_rows, grad_pred,
                                                                                 loss_sum, nullptr, nullptr);
    return after_launch("k_synth_nll");
}

extern "C" int qbold_synth_nll_inferred(const float* labels, int32_t label_stride, const float* pred,
                                        int32_t pred_stride, int32_t use_mvg, const float* inv_gamma_params,
                                        int64_t n, float grad_scale, float* nll_rows, float* grad_pred,
                                        double* loss_sum, double* ig_sums, void* stream) {
    const int nc = use_mvg ? 5 : 4;
    if (n < 0 || label_stride < 2 || pred_stride < nc)
        return fail(QBOLD_EINVAL, "qbold_synth_nll_inferred: bad argument");
    if (n == 0) return QBOLD_OK;
    if (!labels || !pred || !inv_gamma_params || (!nll_rows && !grad_pred && !loss_sum))
        return fail(QBOLD_EINVAL, "qbold_synth_nll_inferred: null pointer");
    SynthOpts opt{};
    opt.use_mvg = use_mvg ? 1 : 0;
    opt.pred_stride = pred_stride;
    opt.inv_gamma = 1;
    k_synth_nll<<<(unsigned)stream_grid(n), kThreads, 0, (cudaStream_t)stream>>>(
        labels, label_stride, pred, opt, n, grad_scale, nll_rows, grad_pred, loss_sum, inv_gamma_params, ig_sums);
    return after_launch("k_synth_nll");
}

extern "C" int qbold_mog_kl(const float* pred, int32_t n_components, const float* mask, const float* eps,
                            uint64_t seed, uint64_t offset, int64_t n, float* kl_map, float* grad_pred,
                            void* stream) {
    if (n < 0 || n_components < 1 || n_components > 64) return fail(QBOLD_EINVAL, "qbold_mog_kl: bad argument");
    if (n == 0) return QBOLD_OK;
    if (!pred || !kl_map) return fail(QBOLD_EINVAL, "qbold_mog_kl: null pointer");
    k_mog_kl<<<(unsigned)stream_grid(n), kThreads, 0, (cudaStream_t)stream>>>(pred, n_components, mask, eps, seed,
                                                                              offset, n, kl_map, grad_pred);
    return after_launch("k_mog_kl");
}

extern "C" int qbold_diag_kl(const float* pred, int32_t pred_stride, const float* prior, int32_t prior_stride,
                             const float* mask, int64_t n, float* kl_map, float* grad_pred, int32_t grad_pred_stride,
                             float* grad_prior, int32_t grad_prior_stride, void* stream) {
    if (n < 0 || pred_stride < 4 || prior_stride < 4 || (grad_pred && grad_pred_stride < 4) ||
        (grad_prior && grad_prior_stride < 4))
        return fail(QBOLD_EINVAL, "qbold_diag_kl: bad argument");
    if (n == 0) return QBOLD_OK;
    if (!pred || !prior || !kl_map) return fail(QBOLD_EINVAL, "qbold_diag_kl: null pointer");
    k_diag_kl<<<(unsigned)stream_grid(n), kThreads, 0, (cudaStream_t)stream>>>(
        pred, pred_stride, prior, prior_stride, mask, n, kl_map, grad_pred, grad_pred_stride, grad_prior,
        grad_prior_stride);
    return after_launch("k_diag_kl");
}

// ---- gated residual mix of the encoder blocks (reference model.py:160-172): out = skip * (1 - g) + r * g with
// g = sigmoid(z + offset).  One streaming pass forward and one backward instead of ~16 elementwise launches.
// z has `zc` channels: C (channel-wise gating) or 1 (one gate per voxel).
namespace qb {

__device__ __forceinline__ float gate_of(float z, float offset) { return 1.0f / (1.0f + expf(-(z + offset))); }

// 16-byte variants for channel-wise gating (same shapes for all operands, total % 4 == 0, aligned pointers)
__global__ void __launch_bounds__(kThreads) k_gate_mix_fwd4(const float4* __restrict__ skip, const float4* __restrict__ r,
                                                            const float4* __restrict__ z, float offset, int64_t total4,
                                                            float4* __restrict__ out) {
    for (int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x; e < total4; e += (int64_t)gridDim.x * kThreads) {
        const float4 zz = __ldg(z + e), s = __ldg(skip + e), rr = __ldg(r + e);
        const float g0 = gate_of(zz.x, offset), g1 = gate_of(zz.y, offset), g2 = gate_of(zz.z, offset),
                    g3 = gate_of(zz.w, offset);
        out[e] = make_float4(s.x * (1.0f - g0) + rr.x * g0, s.y * (1.0f - g1) + rr.y * g1,
                             s.z * (1.0f - g2) + rr.z * g2, s.w * (1.0f - g3) + rr.w * g3);
    }
}

__global__ void __launch_bounds__(kThreads) k_gate_mix_bwd4(const float4* __restrict__ go, const float4* __restrict__ skip,
                                                            const float4* __restrict__ r, const float4* __restrict__ z,
                                                            float offset, int64_t total4, float4* __restrict__ d_skip,
                                                            float4* __restrict__ d_r, float4* __restrict__ d_z) {
    for (int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x; e < total4; e += (int64_t)gridDim.x * kThreads) {
        const float4 zz = __ldg(z + e), s = __ldg(skip + e), rr = __ldg(r + e), o = __ldg(go + e);
        const float g0 = gate_of(zz.x, offset), g1 = gate_of(zz.y, offset), g2 = gate_of(zz.z, offset),
                    g3 = gate_of(zz.w, offset);
        d_skip[e] = make_float4(o.x * (1.0f - g0), o.y * (1.0f - g1), o.z * (1.0f - g2), o.w * (1.0f - g3));
        d_r[e] = make_float4(o.x * g0, o.y * g1, o.z * g2, o.w * g3);
        d_z[e] = make_float4(o.x * (rr.x - s.x) * (g0 * (1.0f - g0)), o.y * (rr.y - s.y) * (g1 * (1.0f - g1)),
                             o.z * (rr.z - s.z) * (g2 * (1.0f - g2)), o.w * (rr.w - s.w) * (g3 * (1.0f - g3)));
    }
}

__global__ void __launch_bounds__(kThreads) k_gate_mix_fwd(const float* __restrict__ skip, const float* __restrict__ r,
                                                           const float* __restrict__ z, float offset, int64_t total,
                                                           int C, int zc, float* __restrict__ out) {
    for (int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x; e < total; e += (int64_t)gridDim.x * kThreads) {
        const float zz = __ldg(z + (zc == 1 ? e / C : e)) + offset;
        const float g = 1.0f / (1.0f + expf(-zz));
        const float s = __ldg(skip + e), rr = __ldg(r + e);
        out[e] = s * (1.0f - g) + rr * g;
    }
}

// d_skip = go (1 - g), d_r = go g, d_z = go (r - skip) g (1 - g)  (summed over channels when zc == 1: C <= 64 lanes
// of consecutive threads hold one voxel, reduced with a segmented loop by the first thread of the voxel)
__global__ void __launch_bounds__(kThreads) k_gate_mix_bwd(const float* __restrict__ go, const float* __restrict__ skip,
                                                           const float* __restrict__ r, const float* __restrict__ z,
                                                           float offset, int64_t total, int C, int zc,
                                                           float* __restrict__ d_skip, float* __restrict__ d_r,
                                                           float* __restrict__ d_z) {
    if (zc == 1) {                                                     // one thread per voxel
        const int64_t nvox = total / C;
        for (int64_t v = (int64_t)blockIdx.x * kThreads + threadIdx.x; v < nvox; v += (int64_t)gridDim.x * kThreads) {
            const float g = 1.0f / (1.0f + expf(-(__ldg(z + v) + offset)));
            float acc = 0.f;
            for (int c = 0; c < C; ++c) {
                const int64_t e = v * C + c;
                const float o = __ldg(go + e), s = __ldg(skip + e), rr = __ldg(r + e);
                d_skip[e] = o * (1.0f - g);
                d_r[e] = o * g;
                acc += o * (rr - s);
            }
            d_z[v] = acc * (g * (1.0f - g));
        }
        return;
    }
    for (int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x; e < total; e += (int64_t)gridDim.x * kThreads) {
        const float g = 1.0f / (1.0f + expf(-(__ldg(z + e) + offset)));
        const float o = __ldg(go + e), s = __ldg(skip + e), rr = __ldg(r + e);
        d_skip[e] = o * (1.0f - g);
        d_r[e] = o * g;
        d_z[e] = o * (rr - s) * (g * (1.0f - g));
    }
}

}  // namespace qb

extern "C" int qbold_gate_mix_forward(const float* skip, const float* r, const float* z, float offset, int64_t n,
                                      int32_t channels, int32_t z_channels, float* out, void* stream) {
    if (n < 0 || channels < 1 || (z_channels != 1 && z_channels != channels))
        return fail(QBOLD_EINVAL, "qbold_gate_mix_forward: bad shape");
    if (n == 0) return QBOLD_OK;
    if (!skip || !r || !z || !out) return fail(QBOLD_EINVAL, "qbold_gate_mix_forward: null pointer");
    const int64_t total = n * channels;
    auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    if (z_channels == channels && (total & 3) == 0 && al(skip) && al(r) && al(z) && al(out)) {
        k_gate_mix_fwd4<<<(unsigned)stream_grid(total / 4), kThreads, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const float4*>(skip), reinterpret_cast<const float4*>(r), reinterpret_cast<const float4*>(z),
            offset, total / 4, reinterpret_cast<float4*>(out));
        return after_launch("k_gate_mix_fwd4");
    }
    k_gate_mix_fwd<<<(unsigned)stream_grid(total), kThreads, 0, (cudaStream_t)stream>>>(skip, r, z, offset, total,
                                                                                        channels, z_channels, out);
    return after_launch("k_gate_mix_fwd");
}

extern "C" int qbold_gate_mix_backward(const float* go, const float* skip, const float* r, const float* z, float offset,
                                       int64_t n, int32_t channels, int32_t z_channels, float* d_skip, float* d_r,
                                       float* d_z, void* stream) {
    if (n < 0 || channels < 1 || (z_channels != 1 && z_channels != channels))
        return fail(QBOLD_EINVAL, "qbold_gate_mix_backward: bad shape");
    if (n == 0) return QBOLD_OK;
    if (!go || !skip || !r || !z || !d_skip || !d_r || !d_z)
        return fail(QBOLD_EINVAL, "qbold_gate_mix_backward: null pointer");
    const int64_t total = n * channels;
    auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    if (z_channels == channels && (total & 3) == 0 && al(go) && al(skip) && al(r) && al(z) && al(d_skip) && al(d_r) &&
        al(d_z)) {
        k_gate_mix_bwd4<<<(unsigned)stream_grid(total / 4), kThreads, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const float4*>(go), reinterpret_cast<const float4*>(skip), reinterpret_cast<const float4*>(r),
            reinterpret_cast<const float4*>(z), offset, total / 4, reinterpret_cast<float4*>(d_skip),
            reinterpret_cast<float4*>(d_r), reinterpret_cast<float4*>(d_z));
        return after_launch("k_gate_mix_bwd4");
    }
    k_gate_mix_bwd<<<(unsigned)stream_grid(z_channels == 1 ? n : total), kThreads, 0, (cudaStream_t)stream>>>(
        go, skip, r, z, offset, total, channels, z_channels, d_skip, d_r, d_z);
    return after_launch("k_gate_mix_bwd");
}
