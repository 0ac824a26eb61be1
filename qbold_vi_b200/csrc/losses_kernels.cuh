// Losses either side of the fused path, each value + analytic gradient in ONE pass:
//   k_smoothness : total-variation term on the encoder output        (reference model.py:726-754)
//   k_synth_nll  : pre-training logit-normal NLL of the labels       (model.py:449-514, 376-421)
//   k_diag_kl    : KL of the diagonal (use_mvg=False) branch         (model.py:685-708)
// All three are HBM-bound elementwise / 4-neighbour stencil kernels: one thread per voxel (row), grid-stride,
// grid = a multiple of the SM count; the neighbour reads of the stencil are served by L1/L2.
//
// The kernels live in this header, apart from their launchers in losses.cu, so that tests/host_emu can compile the same
// kernel source for the host and run it in its SIMT emulator (CPU suite).
#pragma once
#include <math.h>

#include "launch.h"
#include "rng.cuh"

namespace qb {

namespace {

constexpr float kOefRange = 0.8f, kMinOef = 0.04f, kDbvRange = 0.2f, kMinDbv = 0.001f;   // model.py:88-91
constexpr float kExpM2 = 0.1353352832366127f;                                            // np.exp(-2.0), model.py:294
constexpr float kLog2Pi = 1.8378770664093453f;

__device__ __forceinline__ float sigmoidf(float z) { return 1.0f / (1.0f + expf(-z)); }

__device__ __forceinline__ float signf(float d) { return (d > 0.f) ? 1.0f : ((d < 0.f) ? -1.0f : 0.0f); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct Pq {
    float o, d;
};

// forward_transform, then the range rescale of model.py:736-738, written as the reference evaluates it
__device__ __forceinline__ Pq rescaled(const float* __restrict__ q, int64_t v, int n_ch, float& s_o, float& s_d) {
    s_o = sigmoidf(__ldg(q + v * n_ch + 0));
    s_d = sigmoidf(__ldg(q + v * n_ch + 2));
    return Pq{(s_o * kOefRange + kMinOef) / kOefRange, (s_d * kDbvRange + kMinDbv) / kDbvRange};
}

}  // namespace

// One thread per voxel of q [B,X,Y,Z,n_ch].  Value: the edges (x,x+1) and (y,y+1) it owns; gradient: all four
// edges that touch it (sign(0) = 0 like tf.abs; an edge counts only when both ends are inside the mask).
__global__ void __launch_bounds__(kThreads) k_smoothness(const float* __restrict__ q, int n_ch,
                                                         const float* __restrict__ mask, int64_t n, int X, int Y,
                                                         int Z, float scale, const float* __restrict__ scale_dev,
                                                         double* __restrict__ tv_sum, float* __restrict__ grad_q) {
    if (scale_dev != nullptr) scale = __ldg(scale_dev);       // weight / global sum(mask), left on the device
    const int64_t sy = Z, sx = (int64_t)Y * Z;
    double acc = 0.0;
    for (int64_t v = (int64_t)blockIdx.x * kThreads + threadIdx.x; v < n; v += (int64_t)gridDim.x * kThreads) {
        const int64_t r = v / Z;
        const int y = (int)(r % Y), x = (int)((r / Y) % X);
        float g_o = 0.f, g_d = 0.f, s_o = 0.f, s_d = 0.f;
        if (__ldg(mask + v) > 0.f) {
            const Pq p = rescaled(q, v, n_ch, s_o, s_d);
            float t_o, t_d;
            if (x + 1 < X && __ldg(mask + v + sx) > 0.f) {
                const Pq pn = rescaled(q, v + sx, n_ch, t_o, t_d);
                const float d_o = p.o - pn.o, d_d = p.d - pn.d;
                acc += (double)(fabsf(d_o) + fabsf(d_d));
                g_o += signf(d_o);
                g_d += signf(d_d);
            }
            if (y + 1 < Y && __ldg(mask + v + sy) > 0.f) {
                const Pq pn = rescaled(q, v + sy, n_ch, t_o, t_d);
                const float d_o = p.o - pn.o, d_d = p.d - pn.d;
                acc += (double)(fabsf(d_o) + fabsf(d_d));
                g_o += signf(d_o);
                g_d += signf(d_d);
            }
            if (x > 0 && __ldg(mask + v - sx) > 0.f) {
                const Pq pn = rescaled(q, v - sx, n_ch, t_o, t_d);
                g_o -= signf(pn.o - p.o);
                g_d -= signf(pn.d - p.d);
            }
            if (y > 0 && __ldg(mask + v - sy) > 0.f) {
                const Pq pn = rescaled(q, v - sy, n_ch, t_o, t_d);
                g_o -= signf(pn.o - p.o);
                g_d -= signf(pn.d - p.d);
            }
        }
        if (grad_q != nullptr) {
            float* g = grad_q + v * n_ch;
            g[0] = ((g_o * scale) / kOefRange) * kOefRange * (s_o * (1.0f - s_o));
            g[1] = 0.f;
            g[2] = ((g_d * scale) / kDbvRange) * kDbvRange * (s_d * (1.0f - s_d));
            g[3] = 0.f;
            if (n_ch > 4) g[4] = 0.f;
        }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0 && tv_sum != nullptr && acc != 0.0) atomicAdd(tv_sum, acc);
}

struct SynthOpts {
    int use_mvg;
    int inv_gamma;          // 1: subtract the inverse-gamma log-prior of the predicted variances
    float ig_alpha, ig_beta, ig_const;   // ig_const = alpha*log(beta) - lgamma(alpha)
    int pred_stride;        // row stride of pred (>= 5 | 4): the infer_inv_gamma layout carries 4 extra channels
};

// One thread per label row.  labels [n, label_stride >= 2] (OEF, DBV, ...), pred [n, 5 | 4] raw.
__global__ void __launch_bounds__(kThreads) k_synth_nll(const float* __restrict__ labels, int label_stride,
                                                        const float* __restrict__ pred, SynthOpts opt, int64_t n,
                                                        float grad_scale, float* __restrict__ nll_rows,
                                                        float* __restrict__ grad_pred,
                                                        double* __restrict__ loss_sum,
                                                        const float* __restrict__ ig4,
                                                        double* __restrict__ ig_sums) {
    const int nc = opt.use_mvg ? 5 : 4;
    double acc = 0.0;
    // infer_inv_gamma (model.py:493-496): learned (alpha_oef, beta_oef, alpha_dbv, beta_dbv), read from device memory
    float a_o = opt.ig_alpha, b_o = opt.ig_beta, c_o = opt.ig_const, a_d = a_o, b_d = b_o, c_d = c_o;
    float s_lo = 0.f, s_io = 0.f, s_ld = 0.f, s_id = 0.f;     // sums of log v and 1/v: the gradients of the 4 parameters
    if (ig4 != nullptr) {
        a_o = __ldg(ig4 + 0), b_o = __ldg(ig4 + 1), a_d = __ldg(ig4 + 2), b_d = __ldg(ig4 + 3);
        c_o = a_o * logf(b_o) - lgammaf(a_o);
        c_d = a_d * logf(b_d) - lgammaf(a_d);
    }
    for (int64_t v = (int64_t)blockIdx.x * kThreads + threadIdx.x; v < n; v += (int64_t)gridDim.x * kThreads) {
        const float* q = pred + v * opt.pred_stride;
        const float mu_o = __ldg(q + 0), mu_d = __ldg(q + 2);
        const float th1 = tanhf(__ldg(q + 1)), th3 = tanhf(__ldg(q + 3));
        const float ls_o = th1 * 3.0f - 1.0f, ls_d = th3 * 3.0f - 1.0f;              // transform_std, model.py:288-290
        float x_o = (__ldg(labels + v * label_stride + 0) - kMinOef) / kOefRange;    // backwards_transform :307-311
        float x_d = (__ldg(labels + v * label_stride + 1) - kMinDbv) / kDbvRange;
        float th4 = 0.f, raw4 = 0.f, cov = 0.f, konst = 0.f;
        if (opt.use_mvg) {
            x_o = fminf(fmaxf(x_o, 1e-6f), 1.0f - 1e-6f);                            // model.py:394-395
            x_d = fminf(fmaxf(x_d, 1e-6f), 1.0f - 1e-6f);
            raw4 = __ldg(q + 4);
            th4 = tanhf(raw4);
            cov = th4 * kExpM2;
            konst = kLog2Pi;
        }
        const float r_o = logf(x_o / (1.0f - x_o)) - mu_o, r_d = logf(x_d / (1.0f - x_d)) - mu_d;
        const float inv_o = expf(ls_o * -1.0f), inv_d = expf(ls_d * -1.0f);
        const float e_neg = expf(ls_o * -1.0f + ls_d * -1.0f);
        const float inv_bl = (e_neg * cov) * -1.0f;                                  // model.py:434
        const float w_o = r_o * inv_o;
        const float w_d = r_d * inv_d + r_o * inv_bl;
        float loss = konst + 0.5f * (2.0f * (ls_o + ls_d)) + 0.5f * (w_o * w_o + w_d * w_d);
        if (opt.use_mvg)
            loss += (logf(x_o) + logf(1.0f - x_o)) + (logf(x_d) + logf(1.0f - x_d));   // model.py:398
        else
            loss += logf(x_o * (1.0f - x_o)) + logf(x_d * (1.0f - x_d));               // model.py:419
        float d_ls_o = 1.0f - w_o * w_o - w_d * r_o * inv_bl;
        float d_ls_d = 1.0f - w_d * w_d;
        float d_raw4 = (-w_d * r_o * e_neg) * kExpM2 * (1.0f - th4 * th4);
        if (opt.inv_gamma) {                                                         // model.py:495-507
            const float e_o = expf(ls_o), e_d = expf(ls_d);
            const float v_o = opt.use_mvg ? e_o * e_o : expf(ls_o * 2.0f);
            const float vd0 = opt.use_mvg ? e_d * e_d : expf(ls_d * 2.0f);
            const float v_d = opt.use_mvg ? vd0 + raw4 * raw4 : vd0;                 // raw channel 4 (model.py:500)
            const float lv_o = logf(v_o), lv_d = logf(v_d);
            loss -= (c_o - (a_o + 1.0f) * lv_o - b_o / v_o) + (c_d - (a_d + 1.0f) * lv_d - b_d / v_d);
            const float dl_o = (a_o + 1.0f) / v_o - b_o / (v_o * v_o), dl_d = (a_d + 1.0f) / v_d - b_d / (v_d * v_d);
            s_lo += lv_o, s_io += 1.0f / v_o, s_ld += lv_d, s_id += 1.0f / v_d;
            d_ls_o += dl_o * 2.0f * v_o;
            d_ls_d += dl_d * 2.0f * vd0;
            d_raw4 += dl_d * 2.0f * raw4;
        }
        acc += (double)loss;
        if (nll_rows != nullptr) nll_rows[v] = loss;
        if (grad_pred != nullptr) {
            float* g = grad_pred + v * nc;
            g[0] = -(w_o * inv_o + w_d * inv_bl) * grad_scale;
            g[1] = d_ls_o * (3.0f * (1.0f - th1 * th1)) * grad_scale;
            g[2] = -(w_d * inv_d) * grad_scale;
            g[3] = d_ls_d * (3.0f * (1.0f - th3 * th3)) * grad_scale;
            if (opt.use_mvg) g[4] = d_raw4 * grad_scale;
        }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0 && loss_sum != nullptr) atomicAdd(loss_sum, acc);
    if (ig_sums != nullptr) {
        s_lo = warp_sum(s_lo), s_io = warp_sum(s_io), s_ld = warp_sum(s_ld), s_id = warp_sum(s_id);
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(ig_sums + 0, (double)s_lo);
            atomicAdd(ig_sums + 1, (double)s_io);
            atomicAdd(ig_sums + 2, (double)s_ld);
            atomicAdd(ig_sums + 3, (double)s_id);
        }
    }
}

// Mixture-of-Gaussians population prior of kl_loss (reference model.py:666-684), use_mvg=False: a single-sample
// estimate  -entropy(q) + (1/M) sum_i [nll(s_oef; comp_i) + nll(s_dbv; comp_i)]  with s = mean + eps * exp(log_std)
// in logit space and nll(s; m, raw) = ls + 0.5 ((s - m) / exp(ls))^2, ls = transform_std(raw).  pred rows are
// [q (4) | M components (4 each)]; value + analytic gradient w.r.t. all 4 (M + 1) channels (the sample carries the
// reparameterisation gradient).  One thread per voxel, HBM-bound (16 (M+1) B in + 16 (M+1) B out).
__global__ void __launch_bounds__(kThreads) k_mog_kl(const float* __restrict__ pred, int n_comp,
                                                     const float* __restrict__ mask, const float* __restrict__ eps,
                                                     uint64_t seed, uint64_t offset, int64_t n,
                                                     float* __restrict__ kl_map, float* __restrict__ grad_pred) {
    const int width = 4 * (n_comp + 1);
    const float inv_m = 1.0f / (float)n_comp;
    for (int64_t v = (int64_t)blockIdx.x * kThreads + threadIdx.x; v < n; v += (int64_t)gridDim.x * kThreads) {
        const float* q = pred + v * width;
        float* g = grad_pred != nullptr ? grad_pred + v * width : nullptr;
        const bool live = mask == nullptr || __ldg(mask + v) > 0.f;                  // model.py:717
        if (!live) {
            kl_map[v] = 0.f;
            if (g != nullptr)
                for (int c = 0; c < width; ++c) g[c] = 0.f;
            continue;
        }
        float e0, e1;
        if (eps != nullptr) {
            e0 = __ldg(eps + v * 2), e1 = __ldg(eps + v * 2 + 1);
        } else {
            normal_pair(seed, offset + (uint64_t)v, kStreamReparam, e0, e1);
        }
        const float th_o = tanhf(__ldg(q + 1)), th_d = tanhf(__ldg(q + 3));
        const float ls_o = th_o * 3.0f - 1.0f, ls_d = th_d * 3.0f - 1.0f;
        const float sd_o = expf(ls_o), sd_d = expf(ls_d);
        const float s_o = __ldg(q + 0) + e0 * sd_o, s_d = __ldg(q + 2) + e1 * sd_d;  // :670-673
        float kl = (ls_o + ls_d) * -1.0f;                                            // :677
        float gs_o = 0.f, gs_d = 0.f;                                                // d kl / d sample
        for (int i = 0; i < n_comp; ++i) {
            const float* c = q + 4 * (i + 1);
#pragma unroll
            for (int k = 0; k < 4; k += 2) {
                const float th = tanhf(__ldg(c + k + 1));
                const float ls = th * 3.0f - 1.0f;
                const float w = ((k == 0 ? s_o : s_d) - __ldg(c + k)) / expf(ls);
                kl += (ls + 0.5f * (w * w)) * inv_m;                                 // :675-682
                const float dw = w / expf(ls) * inv_m;                               // d/d sample = -d/d mean
                if (k == 0) gs_o += dw;
                else gs_d += dw;
                if (g != nullptr) {
                    g[4 * (i + 1) + k] = -dw;
                    g[4 * (i + 1) + k + 1] = (1.0f - w * w) * inv_m * (3.0f * (1.0f - th * th));
                }
            }
        }
        kl_map[v] = kl;
        if (g != nullptr) {
            g[0] = gs_o;
            g[1] = (gs_o * e0 * sd_o - 1.0f) * (3.0f * (1.0f - th_o * th_o));
            g[2] = gs_d;
            g[3] = (gs_d * e1 * sd_d - 1.0f) * (3.0f * (1.0f - th_d * th_d));
        }
    }
}

// KL(LogitNormal q || LogitNormal p) = KL of the underlying Normals, OEF + DBV (tfp, reference model.py:695-708).
// pred / prior rows are [mean_o, raw_std_o, mean_d, raw_std_d] at arbitrary row strides, so the population-prior
// layout (q and p side by side in one 8-channel tensor, model.py:687-689) needs no copy.
__global__ void __launch_bounds__(kThreads) k_diag_kl(const float* __restrict__ pred, int pred_stride,
                                                      const float* __restrict__ prior, int prior_stride,
                                                      const float* __restrict__ mask, int64_t n,
                                                      float* __restrict__ kl_map, float* __restrict__ grad_pred,
                                                      int gpred_stride, float* __restrict__ grad_prior,
                                                      int gprior_stride) {
    for (int64_t v = (int64_t)blockIdx.x * kThreads + threadIdx.x; v < n; v += (int64_t)gridDim.x * kThreads) {
        const bool live = mask == nullptr || __ldg(mask + v) > 0.f;                  // model.py:717
        float kl = 0.f, gq[4] = {0.f, 0.f, 0.f, 0.f}, gp[4] = {0.f, 0.f, 0.f, 0.f};
        if (live) {
#pragma unroll
            for (int c = 0; c < 4; c += 2) {
                const float thq = tanhf(__ldg(pred + v * pred_stride + c + 1));
                const float thp = tanhf(__ldg(prior + v * prior_stride + c + 1));
                const float lq = thq * 3.0f - 1.0f, lp = thp * 3.0f - 1.0f;
                const float inv_p = expf(-lp);
                const float d = __ldg(pred + v * pred_stride + c) * inv_p - __ldg(prior + v * prior_stride + c) * inv_p;
                const float dl = lq - lp;
                const float em = expm1f(2.0f * dl);
                kl += 0.5f * (d * d) + 0.5f * em - dl;
                gq[c] = d * inv_p;
                gp[c] = -d * inv_p;
                gq[c + 1] = em * (3.0f * (1.0f - thq * thq));
                gp[c + 1] = (-(d * d) - em) * (3.0f * (1.0f - thp * thp));
            }
        }
        kl_map[v] = kl;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (grad_pred != nullptr) grad_pred[v * gpred_stride + c] = gq[c];
            if (grad_prior != nullptr) grad_prior[v * gprior_stride + c] = gp[c];
        }
    }
}

}  // namespace qb
