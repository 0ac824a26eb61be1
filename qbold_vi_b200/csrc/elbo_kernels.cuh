// K2: fused amortized-VI training step of the likelihood side.
//
// One kernel replaces, per voxel (reference file:line):
//   ReparamTrickLayer.call           model.py:21-50    logit-normal reparameterised sample
//   SignalGenerationLayer.call       signals.py:55-114 forward model (+ J1 sweep for the gradient)
//   fine_tune_loss_fn                model.py:527-568  tau=0 normalisation, Gaussian / Student-t NLL, mask
//   kl_loss -> mvg_kl_samples        model.py:654-665, 592-610, 376-447  MC KL(q || prior)
// and the backward pass TensorFlow autodiff runs through all of it (stop_gradient on q inside
// log q, identity gradient through the clip, bessel_j0' = -bessel_j1).  The predicted signal
// never reaches HBM: only grad_q [n,5], grad_sigma [n,n_tau] and the loss partial sums are written.
//
// The kernels live in this header, apart from their launchers in elbo.cu, so that tests/host_emu can compile the same
// kernel source for the host and run it in its SIMT emulator (CPU suite).
#pragma once
#include "qbold_core.cuh"
#include "launch.h"
#include "rng.cuh"

#ifndef QB_ELBO_MIN_BLOCKS
#define QB_ELBO_MIN_BLOCKS 3
#endif

namespace qb {

constexpr float kOefRange = 0.8f, kMinOef = 0.04f, kDbvRange = 0.2f, kMinDbv = 0.001f;   // model.py:88-91
constexpr float kExpM2 = 0.1353352832366127f;                                            // np.exp(-2.0), model.py:294
constexpr float kLog2Pi = 1.8378770664093453f;                                           // model.py:390
constexpr float kLogSqrt2Pi = 0.9189385332046727f;                                       // model.py:561
constexpr float kRoundTripZ = 9.0f;   // sigmoid/logit round trip treated as identity below this |z| (see kl_term)

__device__ __forceinline__ float sigmoidf(float z) { return 1.0f / (1.0f + expf(-z)); }

// Sum over the live tau lanes (lanes >= n_tau hold 0): 4 butterfly steps cover 16 lanes, a 5th only when n_tau > 16.
__device__ __forceinline__ float sum_live(float v, bool wide) {
    if (wide) v += __shfl_xor_sync(kFull, v, 16);
    v += __shfl_xor_sync(kFull, v, 8);
    v += __shfl_xor_sync(kFull, v, 4);
    v += __shfl_xor_sync(kFull, v, 2);
    v += __shfl_xor_sync(kFull, v, 1);
    return v;
}

// Transformed distribution parameters of one voxel (q and prior), every lane holds a copy.
struct Dist {
    float mu_o, mu_d, ls_o, ls_d, cov, inv_o, inv_d, inv_bl;
};
struct QExtra {
    float sd_o, sd_d, dls_o, dls_d, dcov;
};

// A group of W lanes (32: the whole warp, 16: one half of a paired warp) serves one voxel.  Lanes 0..2 of the
// group transform q[1], q[3], q[4]; lanes 3..5 the prior's; lanes 6/7 the off-diagonal inverse factor.  One tanhf
// + two expf per lane instead of 6 + 8 on every lane; shuffles broadcast inside the group.
template <int W = 32>
__device__ __forceinline__ void load_dists(const float* __restrict__ q, const float* __restrict__ prior,
                                           int lane, Dist& dq, QExtra& ex, Dist& dp) {
    const int t = lane & (W - 1), gb = lane & ~(W - 1);
    const int sel = t % 3;
    const int idx = sel == 0 ? 1 : (sel == 1 ? 3 : 4);
    const float* src = (t < 3 || prior == nullptr) ? q : prior;
    const float raw = (t < 6) ? __ldg(src + idx) : 0.f;
    const float th = tanhf(raw);
    const float ls = th * 3.0f - 1.0f;                       // transform_std, model.py:288-290
    const float e = expf(ls);
    const float ie = expf(ls * -1.0f);                       // model.py:432-433
    const float th1 = __shfl_sync(kFull, th, gb + 0), th3 = __shfl_sync(kFull, th, gb + 1);
    const float th4 = __shfl_sync(kFull, th, gb + 2);
    dq.mu_o = __ldg(q + 0);
    dq.mu_d = __ldg(q + 2);
    dq.ls_o = __shfl_sync(kFull, ls, gb + 0);
    dq.ls_d = __shfl_sync(kFull, ls, gb + 1);
    dq.cov = th4 * kExpM2;                                   // transform_offdiag, model.py:292-294
    dq.inv_o = __shfl_sync(kFull, ie, gb + 0);
    dq.inv_d = __shfl_sync(kFull, ie, gb + 1);
    ex.sd_o = __shfl_sync(kFull, e, gb + 0);
    ex.sd_d = __shfl_sync(kFull, e, gb + 1);
    ex.dls_o = 3.0f * (1.0f - th1 * th1);
    ex.dls_d = 3.0f * (1.0f - th3 * th3);
    ex.dcov = kExpM2 * (1.0f - th4 * th4);
    const float pth4 = __shfl_sync(kFull, th, gb + 5);
    dp.ls_o = __shfl_sync(kFull, ls, gb + 3);
    dp.ls_d = __shfl_sync(kFull, ls, gb + 4);
    dp.cov = pth4 * kExpM2;
    dp.inv_o = __shfl_sync(kFull, ie, gb + 3);
    dp.inv_d = __shfl_sync(kFull, ie, gb + 4);
    dp.mu_o = prior ? __ldg(prior + 0) : 0.f;
    dp.mu_d = prior ? __ldg(prior + 2) : 0.f;
    // inv_bl = exp(-ls_o + -ls_d) * cov * -1   (model.py:434); lanes 6 (q) and 7 (prior) of the group
    const float a = (t == 7) ? dp.ls_o : dq.ls_o, b = (t == 7) ? dp.ls_d : dq.ls_d;
    const float c = (t == 7) ? dp.cov : dq.cov;
    const float bl = (expf(a * -1.0f + b * -1.0f) * c) * -1.0f;
    dq.inv_bl = __shfl_sync(kFull, bl, gb + 6);
    dp.inv_bl = __shfl_sync(kFull, bl, gb + 7);
}

struct Sample {
    float s_o, s_d, oef, dbv;
};

__device__ __forceinline__ Sample draw(const Dist& dq, const QExtra& ex, float e0, float e1) {
    Sample r;
    const float z_o = dq.mu_o + e0 * ex.sd_o;                                  // model.py:26-27
    const float z_d = (dq.mu_d + e0 * dq.cov) + e1 * ex.sd_d;                  // model.py:29-31
    r.s_o = sigmoidf(z_o);
    r.s_d = sigmoidf(z_d);
    r.oef = r.s_o * kOefRange + kMinOef;                                       // model.py:302-303
    r.dbv = r.s_d * kDbvRange + kMinDbv;
    return r;
}

// 0.5 * squared whitened residual + log-det part of the logit-MVN NLL and its gradient w.r.t.
// the (logit-space) observation (model.py:385-390, 423-447).  The Jacobian term (model.py:398)
// is identical in log q and log p and cancels in log q - log p, so it is not evaluated.
__device__ __forceinline__ float mvn_nll(const Dist& d, float zh_o, float zh_d, float& g_o, float& g_d) {
    const float r_o = zh_o - d.mu_o, r_d = zh_d - d.mu_d;
    const float w_o = r_o * d.inv_o;
    const float w_d = r_d * d.inv_d + r_o * d.inv_bl;
    g_o = w_o * d.inv_o + w_d * d.inv_bl;
    g_d = w_d * d.inv_d;
    return kLog2Pi + 0.5f * (2.0f * (d.ls_o + d.ls_d)) + 0.5f * (w_o * w_o + w_d * w_d);
}

// -log StudentT(df, 0, sigma).pdf(res) and its partials (model.py:557-559); cold path (optimal.yaml: df = 200).
// Out of line for the same reason as roundtrip_literal below.
__device__ __noinline__ float3 student_t_terms_cold(float logc, float df, float zq, float sg, float inv_sg) {
    const float t = zq * zq / df;
    const float k = (df + 1.0f) / (df + zq * zq);
    return make_float3(-(logc - logf(sg) - 0.5f * (df + 1.0f) * log1pf(t)), k * zq * inv_sg,
                       inv_sg - k * zq * zq * inv_sg);
}
__device__ __forceinline__ void student_t_terms(float logc, float df, float zq, float sg, float inv_sg, float& nll,
                                                    float& d_res, float& d_sg) {
    const float3 r = student_t_terms_cold(logc, df, zq, sg, inv_sg);
    nll = r.x;
    d_res = r.y;
    d_sg = r.z;
}

// The reference's float32 round trip z -> sigmoid -> OEF/DBV -> backwards_transform -> clip -> logit
// (model.py:302-303, 310-311, 394-396) and d zh/d z, evaluated literally.  Cold path (|z| >= kRoundTripZ).
// Kept out of line (scalar arguments, float4 result in registers): the fused kernel is instruction-cache sensitive.
__device__ __noinline__ float4 roundtrip_literal(float z_o, float z_d) {
    const float s_o = sigmoidf(z_o), s_d = sigmoidf(z_d);
    float x_o = ((s_o * kOefRange + kMinOef) - kMinOef) / kOefRange;
    float x_d = ((s_d * kDbvRange + kMinDbv) - kMinDbv) / kDbvRange;
    x_o = fminf(fmaxf(x_o, 1e-6f), 1.0f - 1e-6f);
    x_d = fminf(fmaxf(x_d, 1e-6f), 1.0f - 1e-6f);
    float4 r;
    r.x = logf(x_o / (1.0f - x_o));                                  // zh_o
    r.y = logf(x_d / (1.0f - x_d));                                  // zh_d
    // d zh / d z: logit'(x) * (1/range) * range * sigmoid'(z); the clip passes the gradient (model.py:395)
    r.z = (s_o * (1.0f - s_o)) / (x_o * (1.0f - x_o));
    r.w = (s_d * (1.0f - s_d)) / (x_d * (1.0f - x_d));
    return r;
}

// KL(q || prior) of one voxel and its gradient w.r.t. the raw q parameters; warp-cooperative
// (lanes = samples), every lane returns the same values.
//   n_samples > 0 : the reference's Monte-Carlo estimator mean_s(log q(z_s) - log p(z_s))
//                   (mvg_kl_samples, model.py:592-610) with its path-derivative gradient
//                   (stop_gradient on q inside log q, model.py:596);
//   n_samples == 0: closed-form KL of the two logit-space Gaussians (textbook formula = the
//                   expectation of the estimator; NOT the reference's unused mvg_kl, whose trace
//                   term has a transposed-inverse slip).
struct KlOut {
    float kl;
    float g[5];
};

__device__ __forceinline__ KlOut kl_closed_form(const Dist& dq, const QExtra& ex, const Dist& dp) {
    KlOut o;
    const float m00 = ex.sd_o * dp.inv_o;
    const float m10 = (dq.cov - dp.cov * m00) * dp.inv_d;
    const float m11 = ex.sd_d * dp.inv_d;
    const float d_o = dp.mu_o - dq.mu_o, d_d = dp.mu_d - dq.mu_d;
    const float w_o = d_o * dp.inv_o;
    const float w_d = (d_d - dp.cov * w_o) * dp.inv_d;
    o.kl = 0.5f * ((m00 * m00 + m10 * m10 + m11 * m11) + (w_o * w_o + w_d * w_d) - 2.0f +
                   2.0f * ((dp.ls_o + dp.ls_d) - (dq.ls_o + dq.ls_d)));
    o.g[0] = -w_o * dp.inv_o + w_d * dp.cov * dp.inv_o * dp.inv_d;
    o.g[1] = (m00 * m00 - m10 * dp.cov * m00 * dp.inv_d - 1.0f) * ex.dls_o;
    o.g[2] = -w_d * dp.inv_d;
    o.g[3] = (m11 * m11 - 1.0f) * ex.dls_d;
    o.g[4] = (m10 * dp.inv_d) * ex.dcov;
    return o;
}

template <int W = 32>
__device__ __forceinline__ KlOut kl_term(const Dist& dq, const QExtra& ex, const Dist& dp,
                                         const float* __restrict__ eps_v, uint64_t seed, uint64_t index,
                                         int n_samples, int lane) {
    KlOut o;
    if (n_samples > 0) {
        float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const int t = lane & (W - 1);
        // Rounds of 2W samples: lane t takes the two samples of ONE Philox call (2t, 2t+1); a tail of <= W samples
        // is spread one per lane instead (70 samples: W=16 -> 3 calls + 5 evaluations, W=32 -> 2 + 3).
        for (int base = 0; base < n_samples; base += 2 * W) {
            const int rem = n_samples - base;
            const bool pairwise = rem > W;
            const int s0 = base + (pairwise ? 2 * t : t);
            const int cnt = pairwise ? min(2, n_samples - s0) : (t < rem ? 1 : 0);
            U4 r = {0u, 0u, 0u, 0u};
            if (eps_v == nullptr && cnt > 0) r = mc_words(seed, index, s0);
#pragma unroll 1
            for (int h = 0; h < cnt; ++h) {
                const int sidx = s0 + h;
                float k0, k1;
                if (eps_v) {
                    const float2 e = __ldg(reinterpret_cast<const float2*>(eps_v) + sidx);
                    k0 = e.x;
                    k1 = e.y;
                } else {
                    mc_normal_pair(r, sidx, k0, k1);
                }
                // Sample in logit space, then the reference's round trip sigmoid -> OEF/DBV -> backwards_transform ->
                // clip -> logit (model.py:393-396, 307-316, 10-12).  For |z| < kRoundTripZ the round trip is the
                // identity up to float32 noise (< 5e-4 absolute on zh, zero-mean; far inside the 1e-4 ELBO bar after
                // averaging) and d zh/d z = 1, so it is skipped; beyond it (saturating sigmoid, clip at 1e-6) the
                // reference's float32 arithmetic is followed literally.
                const float z_o = dq.mu_o + k0 * ex.sd_o;                            // model.py:26-27
                const float z_d = (dq.mu_d + k0 * dq.cov) + k1 * ex.sd_d;            // model.py:29-31
                float zh_o = z_o, zh_d = z_d, dz_o = 1.0f, dz_d = 1.0f;
                if (fmaxf(fabsf(z_o), fabsf(z_d)) >= kRoundTripZ) {
                    const float4 rt = roundtrip_literal(z_o, z_d);
                    zh_o = rt.x;
                    zh_d = rt.y;
                    dz_o = rt.z;
                    dz_d = rt.w;
                }
                float gq_o, gq_d, gp_o, gp_d;
                const float nq = mvn_nll(dq, zh_o, zh_d, gq_o, gq_d);
                const float np = mvn_nll(dp, zh_o, zh_d, gp_o, gp_d);
                a[5] += np - nq;                                                     // log q - log p (model.py:603)
                const float hz_o = (gp_o - gq_o) * dz_o, hz_d = (gp_d - gq_d) * dz_d;
                a[0] += hz_o;
                a[1] += hz_o * k0;
                a[2] += hz_d;
                a[3] += hz_d * k1;
                a[4] += hz_d * k0;
            }
        }
        const float inv_s = 1.0f / (float)n_samples;
        if (W == 32) {
            const float tot = butterfly8(a, lane);
            a[0] = __shfl_sync(kFull, tot, butterfly8_src_lane(0));
            a[1] = __shfl_sync(kFull, tot, butterfly8_src_lane(1));
            a[2] = __shfl_sync(kFull, tot, butterfly8_src_lane(2));
            a[3] = __shfl_sync(kFull, tot, butterfly8_src_lane(3));
            a[4] = __shfl_sync(kFull, tot, butterfly8_src_lane(4));
            a[5] = __shfl_sync(kFull, tot, butterfly8_src_lane(5));
        } else {
#pragma unroll
            for (int i = 0; i < 6; ++i) {
#pragma unroll
                for (int o2 = W / 2; o2 > 0; o2 >>= 1) a[i] += __shfl_xor_sync(kFull, a[i], o2);
            }
        }
        o.g[0] = a[0] * inv_s;
        o.g[1] = a[1] * inv_s * ex.sd_o * ex.dls_o;
        o.g[2] = a[2] * inv_s;
        o.g[3] = a[3] * inv_s * ex.sd_d * ex.dls_d;
        o.g[4] = a[4] * inv_s * ex.dcov;
        o.kl = a[5] * inv_s;
    } else {
        o = kl_closed_form(dq, ex, dp);
    }
    return o;
}

// ---- thread-per-voxel KL --------------------------------------------------------------------------------------
// The same estimator as kl_term with ONE LANE per voxel (no shuffles, no idle sample slots, both samples of every
// Philox call used).  While every sample stays inside |z| < kRoundTripZ the round trip is the identity, so
//     zh = mu_q + L k,   log q - log p = C + 1/2 (|b + B k|^2 - |k|^2),   d(log q - log p)/d zh = c + D k
// are polynomials of degree <= 2 in the draw k = (k0, k1): the sums over those samples follow exactly from their five
// moments S0 = sum k0, S1 = sum k1, S00 = sum k0^2, S01 = sum k0 k1, S11 = sum k1^2, and the loop is
// Philox + Box-Muller + 5 FMA + the |z| range check.  A sample at |z| >= kRoundTripZ stays out of the moments and is
// evaluated on its own with the reference's literal float32 round trip (kl_sample_literal, cold).
__device__ __forceinline__ void dists_of_thread(const float* __restrict__ q, const float* __restrict__ prior, Dist& dq,
                                                QExtra& ex, Dist& dp) {
    const float q0 = __ldg(q + 0), q1 = __ldg(q + 1), q2 = __ldg(q + 2), q3 = __ldg(q + 3), q4 = __ldg(q + 4);
    float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f, p4 = 0.f;
    if (prior != nullptr) {
        p0 = __ldg(prior + 0);
        p1 = __ldg(prior + 1);
        p2 = __ldg(prior + 2);
        p3 = __ldg(prior + 3);
        p4 = __ldg(prior + 4);
    }
    const float th1 = tanhf(q1), th3 = tanhf(q3), th4 = tanhf(q4);
    const float ph1 = tanhf(p1), ph3 = tanhf(p3), ph4 = tanhf(p4);
    dq.mu_o = q0;
    dq.mu_d = q2;
    dq.ls_o = th1 * 3.0f - 1.0f;                             // transform_std, model.py:288-290
    dq.ls_d = th3 * 3.0f - 1.0f;
    dq.cov = th4 * kExpM2;                                   // transform_offdiag, model.py:292-294
    dq.inv_o = expf(dq.ls_o * -1.0f);                        // model.py:432-433
    dq.inv_d = expf(dq.ls_d * -1.0f);
    dq.inv_bl = (expf(dq.ls_o * -1.0f + dq.ls_d * -1.0f) * dq.cov) * -1.0f;     // model.py:434
    ex.sd_o = expf(dq.ls_o);
    ex.sd_d = expf(dq.ls_d);
    ex.dls_o = 3.0f * (1.0f - th1 * th1);
    ex.dls_d = 3.0f * (1.0f - th3 * th3);
    ex.dcov = kExpM2 * (1.0f - th4 * th4);
    dp.mu_o = p0;
    dp.mu_d = p2;
    dp.ls_o = ph1 * 3.0f - 1.0f;
    dp.ls_d = ph3 * 3.0f - 1.0f;
    dp.cov = ph4 * kExpM2;
    dp.inv_o = expf(dp.ls_o * -1.0f);
    dp.inv_d = expf(dp.ls_d * -1.0f);
    dp.inv_bl = (expf(dp.ls_o * -1.0f + dp.ls_d * -1.0f) * dp.cov) * -1.0f;
}

// One sample outside |z| < kRoundTripZ, evaluated as in kl_term with the reference's literal float32 round trip and
// added to a[0..5]; cold and out of line (Dist / a[] live on the stack only for this call).
__device__ __noinline__ void kl_sample_literal(const Dist& dq, const QExtra& ex, const Dist& dp, float k0, float k1,
                                               float* __restrict__ a) {
    const float z_o = dq.mu_o + k0 * ex.sd_o;                                // model.py:26-27
    const float z_d = (dq.mu_d + k0 * dq.cov) + k1 * ex.sd_d;                // model.py:29-31
    const float4 rt = roundtrip_literal(z_o, z_d);
    float gq_o, gq_d, gp_o, gp_d;
    const float nq = mvn_nll(dq, rt.x, rt.y, gq_o, gq_d);
    const float np = mvn_nll(dp, rt.x, rt.y, gp_o, gp_d);
    const float hz_o = (gp_o - gq_o) * rt.z, hz_d = (gp_d - gq_d) * rt.w;
    a[0] += hz_o;
    a[1] += hz_o * k0;
    a[2] += hz_d;
    a[3] += hz_d * k1;
    a[4] += hz_d * k0;
    a[5] += np - nq;                                                         // log q - log p (model.py:603)
}

// KL(q || prior) of voxel `index` and its gradient w.r.t. the raw q parameters, computed by the calling lane alone.
// eps_v: this voxel's explicit draws [n_samples, 2] or nullptr (Philox draws of mc_normal_pair).
__device__ __forceinline__ KlOut kl_voxel(const Dist& dq, const QExtra& ex, const Dist& dp,
                                          const float* __restrict__ eps_v, uint64_t seed, uint64_t index,
                                          int n_samples) {
    if (n_samples <= 0) return kl_closed_form(dq, ex, dp);
    float s0 = 0.f, s1 = 0.f, s00 = 0.f, s01 = 0.f, s11 = 0.f;
    float a[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};              // the literally evaluated samples
    int n_in = 0;                                             // samples inside the range: they enter through the moments
#pragma unroll 1
    for (int s = 0; s < n_samples; s += 2) {
        const bool two = s + 1 < n_samples;
        float k0, k1, k2 = 0.f, k3 = 0.f;
        if (eps_v) {
            const float2 e = __ldg(reinterpret_cast<const float2*>(eps_v) + s);
            k0 = e.x;
            k1 = e.y;
            if (two) {
                const float2 e2 = __ldg(reinterpret_cast<const float2*>(eps_v) + s + 1);
                k2 = e2.x;
                k3 = e2.y;
            }
        } else {
            const U4 r = mc_words(seed, index, s);
            mc_box_muller(r.x, r.y, k0, k1);
            mc_box_muller(r.z, r.w, k2, k3);
        }
        const float za = dq.mu_o + k0 * ex.sd_o, zb = (dq.mu_d + k0 * dq.cov) + k1 * ex.sd_d;
        const float zc = dq.mu_o + k2 * ex.sd_o, zd = (dq.mu_d + k2 * dq.cov) + k3 * ex.sd_d;
        if (!(fmaxf(fabsf(za), fabsf(zb)) < kRoundTripZ)) {
            kl_sample_literal(dq, ex, dp, k0, k1, a);
            k0 = k1 = 0.f;
        } else {
            ++n_in;
        }
        if (!two) {
            k2 = k3 = 0.f;
        } else if (!(fmaxf(fabsf(zc), fabsf(zd)) < kRoundTripZ)) {
            kl_sample_literal(dq, ex, dp, k2, k3, a);
            k2 = k3 = 0.f;
        } else {
            ++n_in;
        }
        s0 += k0;
        s1 += k1;
        s00 = fmaf(k0, k0, s00);
        s01 = fmaf(k0, k1, s01);
        s11 = fmaf(k1, k1, s11);
        s0 += k2;
        s1 += k3;
        s00 = fmaf(k2, k2, s00);
        s01 = fmaf(k2, k3, s01);
        s11 = fmaf(k3, k3, s11);
    }
    const float ns = (float)n_samples;
    {
        // the n_in samples inside the range: w_p = b + B k, w_q = k (inverse Cholesky factor of q times its factor)
        const float ni = (float)n_in;
        const float del_o = dq.mu_o - dp.mu_o, del_d = dq.mu_d - dp.mu_d;
        const float b0 = del_o * dp.inv_o, b1 = del_d * dp.inv_d + del_o * dp.inv_bl;
        const float B00 = dp.inv_o * ex.sd_o, B10 = dp.inv_bl * ex.sd_o + dp.inv_d * dq.cov, B11 = dp.inv_d * ex.sd_d;
        const float u0 = B00 * b0 + B10 * b1, u1 = B11 * b1;
        const float sum_wp = ni * (b0 * b0 + b1 * b1) + 2.0f * (u0 * s0 + u1 * s1) + (B00 * B00 + B10 * B10) * s00 +
                             2.0f * (B10 * B11) * s01 + (B11 * B11) * s11;
        a[5] += ni * ((dp.ls_o + dp.ls_d) - (dq.ls_o + dq.ls_d)) + 0.5f * (sum_wp - (s00 + s11));
        // d(log q - log p)/d zh = c + D k
        const float c0 = dp.inv_o * b0 + dp.inv_bl * b1, c1 = dp.inv_d * b1;
        const float D00 = (dp.inv_o * B00 + dp.inv_bl * B10) - dq.inv_o, D01 = dp.inv_bl * B11 - dq.inv_bl;
        const float D10 = dp.inv_d * B10, D11 = dp.inv_d * B11 - dq.inv_d;
        a[0] += ni * c0 + D00 * s0 + D01 * s1;
        a[1] += c0 * s0 + D00 * s00 + D01 * s01;
        a[2] += ni * c1 + D10 * s0 + D11 * s1;
        a[3] += c1 * s1 + D10 * s01 + D11 * s11;
        a[4] += c1 * s0 + D10 * s00 + D11 * s01;
    }
    KlOut o;
    const float inv_s = 1.0f / ns;
    o.g[0] = a[0] * inv_s;
    o.g[1] = a[1] * inv_s * ex.sd_o * ex.dls_o;
    o.g[2] = a[2] * inv_s;
    o.g[3] = a[3] * inv_s * ex.sd_d * ex.dls_d;
    o.g[4] = a[4] * inv_s * ex.dcov;
    o.kl = a[5] * inv_s;
    return o;
}

// kl_loss alone (model.py:654-665): per-voxel KL map and d kl_map[v] / d q[v,:]; one thread per voxel (kl_voxel).
__global__ void __launch_bounds__(kThreads) k_kl(const float* __restrict__ q, const float* __restrict__ prior,
                                                 const float* __restrict__ mask, const float* __restrict__ eps_kl,
                                                 uint64_t seed, uint64_t offset, int n_samples, int64_t n,
                                                 float* __restrict__ kl_map, float* __restrict__ grad_q) {
    for (int64_t v = (int64_t)blockIdx.x * kThreads + threadIdx.x; v < n; v += (int64_t)gridDim.x * kThreads) {
        KlOut ko;
        ko.kl = 0.f;
#pragma unroll
        for (int i = 0; i < 5; ++i) ko.g[i] = 0.f;
        if (mask == nullptr || __ldg(mask + v) > 0.f) {                          // model.py:661
            Dist dq, dp;
            QExtra ex;
            dists_of_thread(q + v * 5, prior + v * 5, dq, ex, dp);
            ko = kl_voxel(dq, ex, dp, eps_kl ? eps_kl + v * n_samples * 2 : nullptr, seed, offset + (uint64_t)v,
                          n_samples);
        }
        kl_map[v] = ko.kl;
        if (grad_q != nullptr) {
#pragma unroll
            for (int i = 0; i < 5; ++i) grad_q[v * 5 + i] = ko.g[i];
        }
    }
}

template <bool HAS_PRIOR, int PATH>
__global__ void __launch_bounds__(kThreads, 3) k_elbo(const __grid_constant__ QboldParams P,
                                                   const float* __restrict__ q, const float* __restrict__ sigma,
                                                   const float* __restrict__ y, const float* __restrict__ mask,
                                                   const float* __restrict__ prior, const float* __restrict__ eps,
                                                   const float* __restrict__ eps_kl, uint64_t seed,
                                                   const uint64_t* __restrict__ seed_dev, uint64_t offset,
                                                   int kl_samples, float inv_mask_sum,
                                                   const float* __restrict__ inv_mask_sum_dev, float kl_weight, int64_t n,
                                                   float* __restrict__ grad_q, float* __restrict__ grad_sigma,
                                                   float* __restrict__ nll_map, float* __restrict__ kl_map,
                                                   double* __restrict__ sums, unsigned long long* __restrict__ work) {
    __shared__ QuadSmem s;
    __shared__ SchedSmem ss;
    if (P.full_model) {
        if (PATH == kSched) load_sched(P, ss);
        else load_quad_tables(P, s);
    }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    const int nt = P.n_tau;
    const bool live = lane < nt;
    const int my_col = live ? P.col_of_tau[lane] : -1;
    const float my_tau = live ? P.tau[lane] : 0.f;
    const float my_b = live ? P.blood_b[lane] : 0.f;
    const int se = P.se_idx;
    const bool multi = P.multi_image_normalisation != 0;
    const bool in_norm = multi ? (lane >= se - 1 && lane <= se + 1) : (lane == se);
    const float norm_w = multi ? (1.0f / 3.0f) : 1.0f;
    const QuadCtx qc = make_quad_ctx<PATH>(P, ss, lane, my_col, my_tau);
    const float df = P.student_t_df;
    if (inv_mask_sum_dev != nullptr) inv_mask_sum = __ldg(inv_mask_sum_dev);   // 1 / global sum(mask), left on the device
    if (seed_dev != nullptr) seed = __ldg(seed_dev);   // Philox key kept on the device: a captured launch replays with fresh draws
    const bool wide = nt > 16;

    double acc_nll = 0.0, acc_kl = 0.0, acc_mask = 0.0;
    int bad = 0;

    for (int64_t v = next_unit(work, lane), nxt_unit; v < n; v = nxt_unit) {   // dynamic units, see next_unit
        nxt_unit = next_unit(work, lane);
        const float m = __ldg(mask + v);
        if (!(m != 0.0f)) {
            // masked voxel: nll * 0 and where(mask > 0, kl, 0) (model.py:564,661) -> zero loss and gradient
            if (lane < 5) grad_q[v * 5 + lane] = 0.f;
            if (live) grad_sigma[v * nt + lane] = 0.f;
            if (lane == 0) {
                if (nll_map) nll_map[v] = 0.f;
                if (kl_map) kl_map[v] = 0.f;
            }
            continue;
        }
        Dist dq, dp;
        QExtra ex;
        load_dists(q + v * 5, HAS_PRIOR ? prior + v * 5 : nullptr, lane, dq, ex, dp);

        // ---- likelihood sample
        float e0, e1;
        if (eps) {
            e0 = __ldg(eps + v * 2);
            e1 = __ldg(eps + v * 2 + 1);
        } else {
            normal_pair(seed, offset + (uint64_t)v, kStreamReparam, e0, e1);
        }
        const Sample sm = draw(dq, ex, e0, e1);
        const VoxelPhys vp = voxel_phys<false>(P, sm.oef, sm.dbv, P.hct);
        float I = 0.f, dI = 0.f;
        if (P.full_model) tissue_eval<true, PATH>(P, s, ss, qc, vp.dw, vp.dw_k, I, dI);
        const TauSignal ts = tau_signal<true>(P, vp, my_tau, my_b, I, dI);

        // ---- fine_tune_loss_fn (model.py:527-568).  Divisions by the same denominator share one reciprocal.
        const float yv = live ? __ldg(y + v * nt + lane) : 0.f;
        const float sg = live ? __ldg(sigma + v * nt + lane) : 1.f;
        const float pred = live ? ts.S : 0.f;
        float npd, ny;                                                          // model.py:541-545
        if (multi) {
            npd = sum_live(in_norm ? pred * norm_w : 0.f, wide) + 1e-3f;
            ny = sum_live(in_norm ? yv * norm_w : 0.f, wide) + 1e-3f;
        } else {
            npd = __shfl_sync(kFull, pred, se) + 1e-3f;
            ny = __shfl_sync(kFull, yv, se) + 1e-3f;
        }
        const float inv_npd = 1.0f / npd, inv_sg = 1.0f / sg;
        float yn = yv / ny, pn = pred * inv_npd;
        float dpn = 1.0f;                                                       // d(pn used in residual)/d(pred/npd)
        if (P.predict_log_data) {                                               // model.py:547-549 (mask > 0 here)
            dpn = 1.0f / pn;
            yn = logf(yn);
            pn = logf(pn);
        }
        const float res = yn - pn;
        const float zq = res * inv_sg;
        float nll_t, dnll_dres, dnll_dsg;
        if (df > 0.f) {                                                         // StudentT(df, 0, sigma), model.py:557-559
            student_t_terms(P.student_t_logc, df, zq, sg, inv_sg, nll_t, dnll_dres, dnll_dsg);
        } else {                                                                // Gaussian, model.py:561
            nll_t = -(-logf(sg) - kLogSqrt2Pi - 0.5f * (zq * zq));
            dnll_dres = zq * inv_sg;
            dnll_dsg = inv_sg - (zq * zq) * inv_sg;
        }
        if (!live) nll_t = 0.f;
        const float scale = m * inv_mask_sum;                                   // model.py:564-566
        if (live) grad_sigma[v * nt + lane] = dnll_dsg * scale;
        // residual = yn - pn  =>  d/dpn = -dnll_dres ;  pn = pred/npd, npd = pred[se] (+ neighbours) + 1e-3:
        //   dL/dpred_t = g_t/npd + [t in norm] * w * g_npd,  g_npd = -sum_t g_t pred_t / npd^2
        const float g_ratio = live ? (-dnll_dres * scale) * dpn : 0.f;          // w.r.t. pred/npd
        const float nll_v = sum_live(nll_t, wide);
        const float s_gp = sum_live(g_ratio * pred, wide);
        const float s_go = sum_live(g_ratio * ts.dS_doef, wide);
        const float s_gd = sum_live(g_ratio * ts.dS_ddbv, wide);
        float n_o, n_d;                                                          // sum over the normalisation set of dS/d.
        if (multi) {
            n_o = sum_live(in_norm ? ts.dS_doef * norm_w : 0.f, wide);
            n_d = sum_live(in_norm ? ts.dS_ddbv * norm_w : 0.f, wide);
        } else {
            n_o = __shfl_sync(kFull, ts.dS_doef, se);
            n_d = __shfl_sync(kFull, ts.dS_ddbv, se);
        }
        const float g_npd = -s_gp * (inv_npd * inv_npd);
        const float go = s_go * inv_npd + g_npd * n_o;
        const float gd = s_gd * inv_npd + g_npd * n_d;
        float gz_o = go * kOefRange * sm.s_o * (1.0f - sm.s_o);
        float gz_d = gd * kDbvRange * sm.s_d * (1.0f - sm.s_d);
        // z -> q (model.py:26-31)
        float g0 = gz_o, g1 = gz_o * e0 * ex.sd_o * ex.dls_o, g2 = gz_d;
        float g3 = gz_d * e1 * ex.sd_d * ex.dls_d, g4 = gz_d * e0 * ex.dcov;

        // ---- KL(q || prior), where(mask > 0) (model.py:661)
        float kl_v = 0.f;
        if (HAS_PRIOR && m > 0.f) {
            const KlOut ko = kl_term(dq, ex, dp, eps_kl ? eps_kl + v * kl_samples * 2 : nullptr, seed,
                                     offset + (uint64_t)v, kl_samples, lane);
            kl_v = ko.kl;
            const float w = kl_weight * inv_mask_sum;
            g0 += w * ko.g[0];
            g1 += w * ko.g[1];
            g2 += w * ko.g[2];
            g3 += w * ko.g[3];
            g4 += w * ko.g[4];
        }

        if (lane < 5) {
            const float gv = lane == 0 ? g0 : lane == 1 ? g1 : lane == 2 ? g2 : lane == 3 ? g3 : g4;
            grad_q[v * 5 + lane] = gv;
        }
        if (lane == 0) {
            const float nm = nll_v * m;
            if (nll_map) nll_map[v] = nm;
            if (kl_map) kl_map[v] = kl_v;
            acc_nll += (double)nm;
            acc_kl += (double)kl_v;
            acc_mask += (double)m;
            if (!isfinite(nm + kl_v)) bad = 1;
        }
    }
    if (lane == 0 && sums != nullptr) {
        if (acc_mask != 0.0 || bad) {
            atomicAdd(sums + 0, acc_nll);
            atomicAdd(sums + 1, acc_kl);
            atomicAdd(sums + 2, acc_mask);
            if (bad) atomicAdd(sums + 3, 1.0);
        }
    }
}

// Paired variant of k_elbo for the scheduled path with n_tau <= 16: a warp takes TWO voxels per iteration.  Only the
// two quadratures run on all 32 lanes (one after the other); the parameter transforms, the sample, the per-tau
// likelihood, its reductions, the KL (lanes = samples, 16 per pass) and all loads / stores are done once for both
// voxels, voxel 0 on lanes 0-15 and voxel 1 on lanes 16-31.
//
// Work unit = 32 consecutive voxels per warp, in two phases.  Phase A, lanes = voxels: every lane does the per-voxel
// scalar work of ONE voxel by itself -- parameter transforms, the reparameterised draw (Philox + accurate Box-Muller),
// sigmoid / OEF / DBV, and the whole KL term (kl_voxel: a compact Philox / Box-Muller / five-moment loop that all lanes
// of the warp run in lock step) -- and parks 14 floats in the warp's shared-memory slot.  Phase B, the 16 pairs of
// the unit: quadrature, likelihood, reductions and stores as described above, fed from the slot.
template <bool HAS_PRIOR>
__global__ void __launch_bounds__(kThreads, QB_ELBO_MIN_BLOCKS) k_elbo_pair(const __grid_constant__ QboldParams P,
                                                           const float* __restrict__ q, const float* __restrict__ sigma,
                                                           const float* __restrict__ y, const float* __restrict__ mask,
                                                           const float* __restrict__ prior, const float* __restrict__ eps,
                                                           const float* __restrict__ eps_kl, uint64_t seed,
                                                           const uint64_t* __restrict__ seed_dev, uint64_t offset,
                                                           int kl_samples, float inv_mask_sum,
                                                           const float* __restrict__ inv_mask_sum_dev,
                                                           float kl_weight, int64_t n, float* __restrict__ grad_q,
                                                           float* __restrict__ grad_sigma, float* __restrict__ nll_map,
                                                           float* __restrict__ kl_map, double* __restrict__ sums,
                                                           unsigned long long* __restrict__ work) {
    __shared__ SchedSmem ss;
    // [warp][voxel of the unit][kl, kl gradient 0..4, oef, dbv, d oef/d z_o, d dbv/d z_d, d z_o/d raw1, d z_d/d raw3,
    //                             d z_d/d raw4, mask, -, -]
    __shared__ __align__(16) float vox_slot[kThreads / 32][32][16];
    load_sched(P, ss);
    __syncthreads();
    const int lane = threadIdx.x & 31, half = lane >> 4, t = lane & 15, gb = lane & 16;
    const int wslot = threadIdx.x >> 5;
    const int nt = P.n_tau;
    const bool live = t < nt;
    const int my_col = live ? P.col_of_tau[t] : -1;
    const float my_tau = live ? P.tau[t] : 0.f;
    const float my_b = live ? P.blood_b[t] : 0.f;
    const int se = P.se_idx;
    const bool multi = P.multi_image_normalisation != 0;
    const bool in_norm = multi ? (t >= se - 1 && t <= se + 1) : (t == se);
    const float norm_w = multi ? (1.0f / 3.0f) : 1.0f;
    const float df = P.student_t_df;
    const QuadCtx qc = make_quad_ctx<kSched>(P, ss, lane, my_col, my_tau);
    if (inv_mask_sum_dev != nullptr) inv_mask_sum = __ldg(inv_mask_sum_dev);   // 1 / global sum(mask), left on the device
    if (seed_dev != nullptr) seed = __ldg(seed_dev);   // Philox key kept on the device: a captured launch replays with fresh draws
    const int64_t npairs = (n + 1) >> 1;
    double acc_nll = 0.0, acc_kl = 0.0, acc_mask = 0.0;
    int bad = 0;

    // units of 16 pairs come from the device work counter (masked volumes: see next_unit); the next index is fetched early
    for (int64_t un = next_unit(work, lane), nxt; un * 16 < npairs; un = nxt) {
      nxt = next_unit(work, lane);
      {
          // ---- phase A: voxel un * 32 + lane on this lane
          const int64_t va = un * 32 + lane;
          const float ma = va < n ? __ldg(mask + va) : 0.f;
          float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = make_float4(0.f, 0.f, 0.4f, 0.05f);
          float4 s2 = make_float4(0.f, 0.f, 0.f, 0.f), s3 = make_float4(0.f, ma, 0.f, 0.f);
          if (ma != 0.0f) {
              Dist dq, dp;
              QExtra ex;
              dists_of_thread(q + va * 5, HAS_PRIOR ? prior + va * 5 : nullptr, dq, ex, dp);
              float e0, e1;
              if (eps) {
                  e0 = __ldg(eps + va * 2);
                  e1 = __ldg(eps + va * 2 + 1);
              } else {
                  normal_pair(seed, offset + (uint64_t)va, kStreamReparam, e0, e1);
              }
              const Sample sm = draw(dq, ex, e0, e1);
              s1.z = sm.oef;
              s1.w = sm.dbv;
              s2.x = kOefRange * sm.s_o * (1.0f - sm.s_o);
              s2.y = kDbvRange * sm.s_d * (1.0f - sm.s_d);
              s2.z = e0 * ex.sd_o * ex.dls_o;
              s2.w = e1 * ex.sd_d * ex.dls_d;
              s3.x = e0 * ex.dcov;
              if (HAS_PRIOR && ma > 0.f) {                   // model.py:661: KL only where mask > 0
                  const KlOut ko = kl_voxel(dq, ex, dp, eps_kl ? eps_kl + va * kl_samples * 2 : nullptr, seed,
                                            offset + (uint64_t)va, kl_samples);
                  s0 = make_float4(ko.kl, ko.g[0], ko.g[1], ko.g[2]);
                  s1.x = ko.g[3];
                  s1.y = ko.g[4];
              }
          }
          __syncwarp();                                      // phase B of the previous unit has read the slot
          float4* slot = reinterpret_cast<float4*>(&vox_slot[wslot][lane][0]);
          slot[0] = s0;
          slot[1] = s1;
          slot[2] = s2;
          slot[3] = s3;
          __syncwarp();
      }
      const int64_t pr_end = (un + 1) * 16 < npairs ? (un + 1) * 16 : npairs;
#pragma unroll 1
      for (int64_t pr = un * 16; pr < pr_end; ++pr) {
        const int64_t v = pr * 2 + half;
        const bool valid = v < n;                            // odd tail: the slot holds mask 0 and harmless OEF / DBV
        const float* slot = &vox_slot[wslot][(int)(pr - un * 16) * 2 + half][0];
        const float m = slot[13];
        const bool on = (m != 0.0f);
        if (!__any_sync(kFull, on)) {
            // both voxels masked: nll * 0 and where(mask > 0, kl, 0) (model.py:564,661) -> zero loss and gradient
            if (valid) {
                if (t < 5) grad_q[v * 5 + t] = 0.f;
                if (live) grad_sigma[v * nt + t] = 0.f;
                if (t == 0) {
                    if (nll_map) nll_map[v] = 0.f;
                    if (kl_map) kl_map[v] = 0.f;
                }
            }
            continue;
        }
        const VoxelPhys vp = voxel_phys<false>(P, slot[6], slot[7], P.hct);
        const float A_mine = qc.tau_ref15 * vp.dw;
        const unsigned on_mask = __ballot_sync(kFull, on);
        float I = 0.f, Dm = 0.f;
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            if (!((on_mask >> (h << 4)) & 1u)) continue;     // this voxel is masked: skip its quadrature
            const float A = __shfl_sync(kFull, A_mine, h << 4);
            float vi, vd;
            tissue_sched<true>(qc.nph, qc.sa, A, lane, qc.ph_lo, qc.ph_hi, my_col, vi, vd);
            if (half == h) {
                I = vi;
                Dm = vd;
            }
        }
        const float a_t = 1.5f * (fabsf(my_tau) * vp.dw);
        const float dI = (Dm * qc.tau_ref15 + qc.node0_d * a_t) * vp.dw_k;
        if (my_col >= 0) I += node0_value(P, a_t);
        const TauSignal ts = tau_signal<true>(P, vp, my_tau, my_b, I, dI);

        // ---- fine_tune_loss_fn (model.py:527-568), all reductions inside the 16-lane half
        const float yv = live ? __ldg(y + v * nt + t) : 0.f;
        const float sg = live ? __ldg(sigma + v * nt + t) : 1.f;
        const float pred = live ? ts.S : 0.f;
        float npd, ny;
        if (multi) {
            npd = sum_live(in_norm ? pred * norm_w : 0.f, false) + 1e-3f;
            ny = sum_live(in_norm ? yv * norm_w : 0.f, false) + 1e-3f;
        } else {
            npd = __shfl_sync(kFull, pred, gb + se) + 1e-3f;
            ny = __shfl_sync(kFull, yv, gb + se) + 1e-3f;
        }
        const float inv_npd = 1.0f / npd, inv_sg = 1.0f / sg;
        float yn = yv / ny, pn = pred * inv_npd;
        float dpn = 1.0f;
        if (P.predict_log_data) {
            dpn = 1.0f / pn;
            yn = logf(yn);
            pn = logf(pn);
        }
        const float zq = (yn - pn) * inv_sg;
        float nll_t, dnll_dres, dnll_dsg;
        if (df > 0.f) {
            student_t_terms(P.student_t_logc, df, zq, sg, inv_sg, nll_t, dnll_dres, dnll_dsg);
        } else {
            nll_t = -(-logf(sg) - kLogSqrt2Pi - 0.5f * (zq * zq));
            dnll_dres = zq * inv_sg;
            dnll_dsg = inv_sg - (zq * zq) * inv_sg;
        }
        if (!live) nll_t = 0.f;
        const float scale = m * inv_mask_sum;
        const float g_ratio = live ? (-dnll_dres * scale) * dpn : 0.f;
        const float nll_v = sum_live(nll_t, false);
        const float s_gp = sum_live(g_ratio * pred, false);
        const float s_go = sum_live(g_ratio * ts.dS_doef, false);
        const float s_gd = sum_live(g_ratio * ts.dS_ddbv, false);
        float n_o, n_d;
        if (multi) {
            n_o = sum_live(in_norm ? ts.dS_doef * norm_w : 0.f, false);
            n_d = sum_live(in_norm ? ts.dS_ddbv * norm_w : 0.f, false);
        } else {
            n_o = __shfl_sync(kFull, ts.dS_doef, gb + se);
            n_d = __shfl_sync(kFull, ts.dS_ddbv, gb + se);
        }
        const float g_npd = -s_gp * (inv_npd * inv_npd);
        const float go = s_go * inv_npd + g_npd * n_o;
        const float gd = s_gd * inv_npd + g_npd * n_d;
        const float gz_o = go * slot[8], gz_d = gd * slot[9];
        // gradient w.r.t. raw q[t], t < 5: likelihood part through the sample + KL part left by phase A
        float g_mine = 0.f;
        if (t < 5) {
            const float gz = t < 2 ? gz_o : gz_d;
            g_mine = (t == 0 || t == 2) ? gz : gz * slot[9 + t - (t > 1)];   // t = 1, 3, 4 -> slot[10], [11], [12]
            if (HAS_PRIOR) g_mine += (kl_weight * inv_mask_sum) * slot[1 + t];
        }
        float kl_v = HAS_PRIOR ? slot[0] : 0.f;
        if (valid) {
            if (!on) g_mine = 0.f;
            if (live) grad_sigma[v * nt + t] = on ? dnll_dsg * scale : 0.f;
            if (t < 5) grad_q[v * 5 + t] = g_mine;
            if (t == 0) {
                const float nm = on ? nll_v * m : 0.f;
                if (!on) kl_v = 0.f;
                if (nll_map) nll_map[v] = nm;
                if (kl_map) kl_map[v] = kl_v;
                acc_nll += (double)nm;
                acc_kl += (double)kl_v;
                acc_mask += (double)m;
                if (!isfinite(nm + kl_v)) bad = 1;
            }
        }
      }
    }
    if (t == 0 && sums != nullptr && (acc_mask != 0.0 || bad)) {
        atomicAdd(sums + 0, acc_nll);
        atomicAdd(sums + 1, acc_kl);
        atomicAdd(sums + 2, acc_mask);
        if (bad) atomicAdd(sums + 3, 1.0);
    }
}

// Posterior-predictive likelihood map (save_predictions, model.py:808-817): the reference averages
// fine_tune_loss_fn(return_mean=False) over 100 stochastic forward passes of the fine-tuner.  Here: one warp per
// voxel loops over n_samples reparameterised draws, runs the forward-only quadrature for each and averages the
// masked per-voxel NLL; nothing but the [n] map is written.
template <int PATH>
__global__ void __launch_bounds__(kThreads, 3) k_nll_map(const __grid_constant__ QboldParams P,
                                                         const float* __restrict__ q, const float* __restrict__ sigma,
                                                         const float* __restrict__ y, const float* __restrict__ mask,
                                                         const float* __restrict__ eps, uint64_t seed, uint64_t offset,
                                                         int n_samples, int64_t n, float* __restrict__ nll_map,
                                                         unsigned long long* __restrict__ work) {
    __shared__ QuadSmem s;
    __shared__ SchedSmem ss;
    if (P.full_model) {
        if (PATH == kSched) load_sched(P, ss);
        else load_quad_tables(P, s);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    const int nt = P.n_tau;
    const bool live = lane < nt;
    const int my_col = live ? P.col_of_tau[lane] : -1;
    const float my_tau = live ? P.tau[lane] : 0.f;
    const float my_b = live ? P.blood_b[lane] : 0.f;
    const int se = P.se_idx;
    const bool multi = P.multi_image_normalisation != 0;
    const bool in_norm = multi ? (lane >= se - 1 && lane <= se + 1) : (lane == se);
    const float norm_w = multi ? (1.0f / 3.0f) : 1.0f;
    const bool wide = nt > 16;
    const float df = P.student_t_df;
    const QuadCtx qc = make_quad_ctx<PATH>(P, ss, lane, my_col, my_tau);
    for (int64_t v = next_unit(work, lane), nxt_unit; v < n; v = nxt_unit) {   // dynamic units, see next_unit
        nxt_unit = next_unit(work, lane);
        const float m = mask ? __ldg(mask + v) : 1.0f;
        if (!(m != 0.0f)) {
            if (lane == 0) nll_map[v] = 0.f;
            continue;
        }
        Dist dq, dp;
        QExtra ex;
        load_dists(q + v * 5, nullptr, lane, dq, ex, dp);
        const float yv = live ? __ldg(y + v * nt + lane) : 0.f;
        const float sg = live ? __ldg(sigma + v * nt + lane) : 1.f;
        const float ny = (multi ? sum_live(in_norm ? yv * norm_w : 0.f, wide) : __shfl_sync(kFull, yv, se)) + 1e-3f;
        float yn = yv / ny;
        if (P.predict_log_data) yn = logf(yn);
        const float inv_sg = 1.0f / sg, log_sg = logf(sg);
        float acc = 0.f;
        for (int sidx = 0; sidx < n_samples; ++sidx) {
            float e0, e1;
            if (eps) {
                const float2 e = __ldg(reinterpret_cast<const float2*>(eps) + (v * n_samples + sidx));
                e0 = e.x;
                e1 = e.y;
            } else {
                mc_normal_pair(seed, offset + (uint64_t)v, sidx, e0, e1);
            }
            const Sample sm = draw(dq, ex, e0, e1);
            const VoxelPhys vp = voxel_phys<false>(P, sm.oef, sm.dbv, P.hct);
            float I = 0.f, dI = 0.f;
            if (P.full_model) tissue_eval<false, PATH>(P, s, ss, qc, vp.dw, vp.dw_k, I, dI);
            const TauSignal ts = tau_signal<false>(P, vp, my_tau, my_b, I, dI);
            const float pred = live ? ts.S : 0.f;
            const float npd = (multi ? sum_live(in_norm ? pred * norm_w : 0.f, wide) : __shfl_sync(kFull, pred, se)) + 1e-3f;
            float pn = pred / npd;
            if (P.predict_log_data) pn = logf(pn);
            const float zq = (yn - pn) * inv_sg;
            float nll_t;
            if (df > 0.f) nll_t = -(P.student_t_logc - log_sg - 0.5f * (df + 1.0f) * log1pf(zq * zq / df));
            else nll_t = -(-log_sg - kLogSqrt2Pi - 0.5f * (zq * zq));
            acc += live ? nll_t : 0.f;
        }
        const float tot = sum_live(acc, wide);
        if (lane == 0) nll_map[v] = (tot / (float)n_samples) * m;
    }
}

// Paired variant of k_nll_map (scheduled path, n_tau <= 16): one voxel per warp, TWO posterior samples per
// iteration -- sample s on lanes 0-15, sample s+1 on lanes 16-31 for everything but the two quadratures.
__global__ void __launch_bounds__(kThreads, 4) k_nll_map_pair(const __grid_constant__ QboldParams P,
                                                              const float* __restrict__ q, const float* __restrict__ sigma,
                                                              const float* __restrict__ y, const float* __restrict__ mask,
                                                              const float* __restrict__ eps, uint64_t seed,
                                                              uint64_t offset, int n_samples, int64_t n,
                                                              float* __restrict__ nll_map,
                                                              unsigned long long* __restrict__ work) {
    __shared__ SchedSmem ss;
    load_sched(P, ss);
    __syncthreads();
    const int lane = threadIdx.x & 31, half = lane >> 4, t = lane & 15, gb = lane & 16;
    const int nt = P.n_tau;
    const bool live = t < nt;
    const int my_col = live ? P.col_of_tau[t] : -1;
    const float my_tau = live ? P.tau[t] : 0.f;
    const float my_b = live ? P.blood_b[t] : 0.f;
    const int se = P.se_idx;
    const bool multi = P.multi_image_normalisation != 0;
    const bool in_norm = multi ? (t >= se - 1 && t <= se + 1) : (t == se);
    const float norm_w = multi ? (1.0f / 3.0f) : 1.0f;
    const float df = P.student_t_df;
    const QuadCtx qc = make_quad_ctx<kSched>(P, ss, lane, my_col, my_tau);
    // one voxel (n_samples forward passes) per grab; the next index is fetched before the current voxel is processed
    for (int64_t v = next_unit(work, lane), nxt; v < n; v = nxt) {
        nxt = next_unit(work, lane);
        const float m = mask ? __ldg(mask + v) : 1.0f;
        if (!(m != 0.0f)) {
            if (lane == 0) nll_map[v] = 0.f;
            continue;
        }
        Dist dq, dp;
        QExtra ex;
        load_dists(q + v * 5, nullptr, lane, dq, ex, dp);
        const float yv = live ? __ldg(y + v * nt + t) : 0.f;
        const float sg = live ? __ldg(sigma + v * nt + t) : 1.f;
        const float ny = (multi ? sum_live(in_norm ? yv * norm_w : 0.f, false) : __shfl_sync(kFull, yv, gb + se)) + 1e-3f;
        float yn = yv / ny;
        if (P.predict_log_data) yn = logf(yn);
        const float inv_sg = 1.0f / sg, log_sg = logf(sg);
        float acc = 0.f;
        // Blocks of 64 samples: every lane first draws and transforms ITS two samples (one Philox call, two
        // Box-Muller pairs, four sigmoids -- once per voxel instead of once per lane per sample); the iterations then
        // only fetch (OEF, DBV) of their sample from the owning lane.
        for (int sb = 0; sb < n_samples; sb += 64) {
            const int mine = sb + 2 * lane;
            float o0 = 0.f, d0 = 0.f, o1 = 0.f, d1 = 0.f;
            if (mine < n_samples) {
                float e0, e1, f0 = 0.f, f1 = 0.f;
                const bool second = mine + 1 < n_samples;
                if (eps) {
                    const float2* ev = reinterpret_cast<const float2*>(eps) + (v * n_samples + mine);
                    const float2 a = __ldg(ev);
                    e0 = a.x;
                    e1 = a.y;
                    if (second) {
                        const float2 b = __ldg(ev + 1);
                        f0 = b.x;
                        f1 = b.y;
                    }
                } else {
                    const U4 r = mc_words(seed, offset + (uint64_t)v, mine);
                    mc_box_muller(r.x, r.y, e0, e1);
                    mc_box_muller(r.z, r.w, f0, f1);
                }
                const Sample a = draw(dq, ex, e0, e1), b = draw(dq, ex, f0, f1);
                o0 = a.oef;
                d0 = a.dbv;
                o1 = b.oef;
                d1 = b.dbv;
            }
            const int cnt = min(64, n_samples - sb);
            for (int s0 = 0; s0 < cnt; s0 += 2) {
                const bool both = s0 + 1 < cnt;
                const bool valid = (half == 0) || both;
                const bool odd = half && both;
                const int src = s0 >> 1;                                  // the lane that owns samples s0, s0 + 1
                const float oa = __shfl_sync(kFull, o0, src), ob = __shfl_sync(kFull, o1, src);
                const float da = __shfl_sync(kFull, d0, src), db = __shfl_sync(kFull, d1, src);
                const float oef = odd ? ob : oa, dbv = odd ? db : da;
                const VoxelPhys vp = voxel_phys<false>(P, oef, dbv, P.hct);
                const float A_mine = qc.tau_ref15 * vp.dw;
                float I = 0.f;
#pragma unroll 1
                for (int h = 0; h < 2; ++h) {
                    if (h == 1 && !both) continue;
                    const float A = __shfl_sync(kFull, A_mine, h << 4);
                    float vi, vd;
                    tissue_sched<false>(qc.nph, qc.sa, A, lane, qc.ph_lo, qc.ph_hi, my_col, vi, vd);
                    if (half == h) I = vi;
                }
                if (my_col >= 0) I += node0_value(P, 1.5f * (fabsf(my_tau) * vp.dw));
                const TauSignal ts = tau_signal<false>(P, vp, my_tau, my_b, I, 0.f);
                const float pred = live ? ts.S : 0.f;
                const float npd = (multi ? sum_live(in_norm ? pred * norm_w : 0.f, false)
                                         : __shfl_sync(kFull, pred, gb + se)) + 1e-3f;
                float pn = pred / npd;
                if (P.predict_log_data) pn = logf(pn);
                const float zq = (yn - pn) * inv_sg;
                float nll_t;
                if (df > 0.f) nll_t = -(P.student_t_logc - log_sg - 0.5f * (df + 1.0f) * log1pf(zq * zq / df));
                else nll_t = -(-log_sg - kLogSqrt2Pi - 0.5f * (zq * zq));
                acc += (live && valid) ? nll_t : 0.f;
            }
        }
        const float tot = warp_sum(acc);
        if (lane == 0) nll_map[v] = (tot / (float)n_samples) * m;
    }
}

// fine_tune_loss_fn alone (model.py:527-568), for predictions that already exist in HBM (the unfused graph of
// build_fine_tuner): per-voxel masked NLL and its partial derivatives w.r.t. the predicted images and sigmas.
// W lanes serve one voxel (W = 16: two voxels per warp when n_tau <= 16).  HBM-bound: 3 reads + 2 writes of [n,n_tau].
template <int W>
__global__ void __launch_bounds__(kThreads) k_nll(const __grid_constant__ QboldParams P, const float* __restrict__ y,
                                                  const float* __restrict__ pred_in, const float* __restrict__ sigma,
                                                  const float* __restrict__ mask, int64_t n, float* __restrict__ nll_map,
                                                  float* __restrict__ d_pred, float* __restrict__ d_sigma) {
    const int lane = threadIdx.x & 31, t = lane & (W - 1), gb = lane & ~(W - 1);
    constexpr int kPer = 32 / W;
    const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    const int nt = P.n_tau;
    const bool live = t < nt;
    const int se = P.se_idx;
    const bool multi = P.multi_image_normalisation != 0;
    const bool in_norm = multi ? (t >= se - 1 && t <= se + 1) : (t == se);
    const float norm_w = multi ? (1.0f / 3.0f) : 1.0f;
    const bool wide = (W == 32) && nt > 16;
    const float df = P.student_t_df;
    const int64_t ngroups = (n + kPer - 1) / kPer;
    for (int64_t gidx = warp; gidx < ngroups; gidx += nwarps) {
        int64_t v = gidx * kPer + (lane / W);
        const bool valid = v < n;
        if (!valid) v = n - 1;
        const float m = mask ? __ldg(mask + v) : 1.0f;
        const float yv = live ? __ldg(y + v * nt + t) : 0.f;
        const float pred = live ? __ldg(pred_in + v * nt + t) : 0.f;
        const float sg = live ? __ldg(sigma + v * nt + t) : 1.f;
        float npd, ny;
        if (multi) {
            npd = sum_live(in_norm ? pred * norm_w : 0.f, wide) + 1e-3f;
            ny = sum_live(in_norm ? yv * norm_w : 0.f, wide) + 1e-3f;
        } else {
            npd = __shfl_sync(kFull, pred, gb + se) + 1e-3f;
            ny = __shfl_sync(kFull, yv, gb + se) + 1e-3f;
        }
        const float inv_npd = 1.0f / npd, inv_sg = 1.0f / sg;
        float yn = yv / ny, pn = pred * inv_npd, dpn = 1.0f;
        if (P.predict_log_data) {                                   // where(mask > 0, log(.), 0), model.py:547-549
            if (m > 0.f) {
                dpn = 1.0f / pn;
                yn = logf(yn);
                pn = logf(pn);
            } else {
                yn = pn = dpn = 0.f;
            }
        }
        const float zq = (yn - pn) * inv_sg;
        float nll_t, dnll_dres, dnll_dsg;
        if (df > 0.f) {
            student_t_terms(P.student_t_logc, df, zq, sg, inv_sg, nll_t, dnll_dres, dnll_dsg);
        } else {
            nll_t = -(-logf(sg) - kLogSqrt2Pi - 0.5f * (zq * zq));
            dnll_dres = zq * inv_sg;
            dnll_dsg = inv_sg - (zq * zq) * inv_sg;
        }
        if (!live) nll_t = 0.f;
        const float g_ratio = live ? (-dnll_dres * m) * dpn : 0.f;
        const float nll_v = sum_live(nll_t, wide);
        const float s_gp = sum_live(g_ratio * pred, wide);
        const float g_npd = -s_gp * (inv_npd * inv_npd);
        if (valid) {
            if (live) {
                if (d_pred) d_pred[v * nt + t] = g_ratio * inv_npd + (in_norm ? g_npd * norm_w : 0.f);
                if (d_sigma) d_sigma[v * nt + t] = dnll_dsg * m;
            }
            if (t == 0) nll_map[v] = nll_v * m;
        }
    }
}

// ReparamTrickLayer alone (model.py:21-50): one thread per voxel.
__global__ void __launch_bounds__(kThreads) k_reparam(const float* __restrict__ q, const float* __restrict__ eps,
                                                      uint64_t seed, uint64_t offset, int64_t n,
                                                      float* __restrict__ out) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const float* qv = q + v * 5;
    float e0, e1;
    if (eps) {
        e0 = eps[v * 2];
        e1 = eps[v * 2 + 1];
    } else {
        normal_pair(seed, offset + (uint64_t)v, kStreamReparam, e0, e1);
    }
    const float sd_o = expf(tanhf(qv[1]) * 3.0f - 1.0f), sd_d = expf(tanhf(qv[3]) * 3.0f - 1.0f);
    const float cov = tanhf(qv[4]) * kExpM2;
    const float z_o = qv[0] + e0 * sd_o;
    const float z_d = (qv[2] + e0 * cov) + e1 * sd_d;
    *reinterpret_cast<float2*>(out + v * 2) =
        make_float2(sigmoidf(z_o) * kOefRange + kMinOef, sigmoidf(z_d) * kDbvRange + kMinDbv);
}

// calculate_means(include_r2p=True, return_stds=True) (model.py:326-343): one THREAD per voxel loops over the samples
// (both samples of every Philox call used, no shuffles).  The reference's two-pass mean / mean((s - mean)^2) is formed
// in one pass around a pivot (the voxel's first sample): var = mean(d^2) - mean(d)^2 with d = s - pivot, which keeps
// the cancellation at the level of the two-pass form since |mean(d)| is of the order of the standard deviation.
__global__ void __launch_bounds__(kThreads) k_posterior_stats(const float dw_k, const float* __restrict__ q,
                                                              const float* __restrict__ eps, uint64_t seed,
                                                              uint64_t offset, int n_samples, int64_t n,
                                                              float* __restrict__ mean3, float* __restrict__ var3) {
    const float inv_n = 1.0f / (float)n_samples;
    for (int64_t v = (int64_t)blockIdx.x * kThreads + threadIdx.x; v < n; v += (int64_t)gridDim.x * kThreads) {
        Dist dq, dp;
        QExtra ex;
        dists_of_thread(q + v * 5, nullptr, dq, ex, dp);
        float p_o = 0.f, p_d = 0.f, p_r = 0.f;                                   // pivots
        float a_o = 0.f, a_d = 0.f, a_r = 0.f, b_o = 0.f, b_d = 0.f, b_r = 0.f;
#pragma unroll 1
        for (int s = 0; s < n_samples; s += 2) {
            float k0, k1, k2 = 0.f, k3 = 0.f;
            const bool two = s + 1 < n_samples;
            if (eps) {
                const float2 e = __ldg(reinterpret_cast<const float2*>(eps) + (v * n_samples + s));
                k0 = e.x;
                k1 = e.y;
                if (two) {
                    const float2 e2 = __ldg(reinterpret_cast<const float2*>(eps) + (v * n_samples + s + 1));
                    k2 = e2.x;
                    k3 = e2.y;
                }
            } else {
                const U4 r = mc_words(seed, offset + (uint64_t)v, s);
                mc_box_muller(r.x, r.y, k0, k1);
                mc_box_muller(r.z, r.w, k2, k3);
            }
            const Sample sa = draw(dq, ex, k0, k1);
            const float ra = (dw_k * sa.oef) * sa.dbv;                           // model.py:516-525
            if (s == 0) {
                p_o = sa.oef;
                p_d = sa.dbv;
                p_r = ra;
            }
            float d = sa.oef - p_o;
            a_o += d;
            b_o = fmaf(d, d, b_o);
            d = sa.dbv - p_d;
            a_d += d;
            b_d = fmaf(d, d, b_d);
            d = ra - p_r;
            a_r += d;
            b_r = fmaf(d, d, b_r);
            if (two) {
                const Sample sb = draw(dq, ex, k2, k3);
                const float rb = (dw_k * sb.oef) * sb.dbv;
                d = sb.oef - p_o;
                a_o += d;
                b_o = fmaf(d, d, b_o);
                d = sb.dbv - p_d;
                a_d += d;
                b_d = fmaf(d, d, b_d);
                d = rb - p_r;
                a_r += d;
                b_r = fmaf(d, d, b_r);
            }
        }
        const float m_o = a_o * inv_n, m_d = a_d * inv_n, m_r = a_r * inv_n;
        mean3[v * 3 + 0] = p_o + m_o;
        mean3[v * 3 + 1] = p_d + m_d;
        mean3[v * 3 + 2] = p_r + m_r;
        var3[v * 3 + 0] = fmaxf(b_o * inv_n - m_o * m_o, 0.f);
        var3[v * 3 + 1] = fmaxf(b_d * inv_n - m_d * m_d, 0.f);
        var3[v * 3 + 2] = fmaxf(b_r * inv_n - m_r * m_r, 0.f);
    }
}

}  // namespace qb
