// Kernels of K1 / K1b (forward signal model, forward + vector-Jacobian product) and of the misalignment augmentation.
// Kept in a header, apart from their launchers in forward.cu, so that tests/host_emu can compile the very same kernel
// source for the host and run it in its SIMT emulator (CPU suite).
#pragma once
#include "qbold_core.cuh"
#include "launch.h"
#include "rng.cuh"

#ifndef QB_FWD_MIN_BLOCKS
#define QB_FWD_MIN_BLOCKS 5
#endif

namespace qb {

// One warp per voxel (grid-stride).  BWD: also g_oef_dbv[n,2]; HCT: oef_dbv rows are (OEF,DBV,Hct).
template <bool BWD, bool HCT, int PATH>
__global__ void __launch_bounds__(kThreads, QB_FWD_MIN_BLOCKS) k_forward(const __grid_constant__ QboldParams P,
                                                      const float* __restrict__ oef_dbv,
                                                      const float* __restrict__ g_signal,
                                                      float* __restrict__ signal,
                                                      float* __restrict__ g_oef_dbv, int64_t n,
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
    constexpr int W = HCT ? 3 : 2;
    const QuadCtx qc = make_quad_ctx<PATH>(P, ss, lane, my_col, my_tau);

    for (int64_t v = next_unit(work, lane), nxt_unit; v < n; v = nxt_unit) {   // dynamic units, see next_unit
        nxt_unit = next_unit(work, lane);
        const float oef = __ldg(oef_dbv + v * W);
        const float dbv = __ldg(oef_dbv + v * W + 1);
        const float hct = HCT ? __ldg(oef_dbv + v * W + 2) : P.hct;
        float gs = 1.0f;
        if (BWD && g_signal != nullptr && live) gs = __ldg(g_signal + v * nt + lane);
        const VoxelPhys vp = voxel_phys<HCT>(P, oef, dbv, hct);

        float I = 0.f, dI = 0.f;
        if (P.full_model) tissue_eval<BWD, PATH>(P, s, ss, qc, vp.dw, vp.dw_k, I, dI);
        const TauSignal ts = tau_signal<BWD>(P, vp, my_tau, my_b, I, dI);
        if (live && signal != nullptr) signal[v * nt + lane] = ts.S;
        if (BWD) {
            float go = live ? gs * ts.dS_doef : 0.f;
            float gd = live ? gs * ts.dS_ddbv : 0.f;
            go = warp_sum(go);
            gd = warp_sum(gd);
            if (HCT) {
                const float gh = warp_sum(live ? gs * ts.dS_dhct : 0.f);
                if (lane < 3) g_oef_dbv[v * 3 + lane] = lane == 0 ? go : (lane == 1 ? gd : gh);
            } else if (lane == 0) {
                *reinterpret_cast<float2*>(g_oef_dbv + v * 2) = make_float2(go, gd);
            }
        }
    }
}

// Misalignment augmentation (reference signals.py:80-96).  A Bernoulli(prob) subset of the voxels is "misaligned":
// the images after a random index in [4, n_tau - 1) see perturbed parameters, OEF + N(0, 0.15) clipped to
// [0.05, 0.8] and DBV + N(0, 0.05) clipped to [0.002, 0.3].  The reference makes OEF/DBV per-image tensors and runs
// the whole forward model on them; every image's signal depends on its own parameters only, so this kernel
// recomputes the forward model of the selected voxels with the perturbed pair and overwrites the late images of
// `signal` (which holds the unperturbed forward model).  Voxels come from the work counter: the ~90 % that are not
// selected cost one Philox call.  Draws: explicit arrays (parity tests feed the reference's recorded draws) or
// Philox call (seed, global voxel index, kStreamMisalign): x -> selection, y -> index, (z, w) -> the two normals.
template <bool HCT, int PATH>
__global__ void __launch_bounds__(kThreads, QB_FWD_MIN_BLOCKS) k_misalign(
    const __grid_constant__ QboldParams P, const float* __restrict__ oef_dbv, int64_t n, float prob,
    const float* __restrict__ sel_u01, const int32_t* __restrict__ from_index, const float* __restrict__ eps,
    uint64_t seed, uint64_t offset, float* __restrict__ signal, unsigned long long* __restrict__ work) {
    __shared__ QuadSmem s;
    __shared__ SchedSmem ss;
    if (P.full_model) {
        if (PATH == kSched) load_sched(P, ss);
        else load_quad_tables(P, s);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int nt = P.n_tau;
    const bool live = lane < nt;
    const int my_col = live ? P.col_of_tau[lane] : -1;
    const float my_tau = live ? P.tau[lane] : 0.f;
    const float my_b = live ? P.blood_b[lane] : 0.f;
    constexpr int W = HCT ? 3 : 2;
    const QuadCtx qc = make_quad_ctx<PATH>(P, ss, lane, my_col, my_tau);
    const int span = nt - 1 - 4;                               // tf.random.uniform(minval=4, maxval=n_tau-1, int32)

    for (int64_t v = next_unit(work, lane), nxt_unit; v < n; v = nxt_unit) {
        nxt_unit = next_unit(work, lane);
        float u, e0, e1;
        int idx;
        if (sel_u01 != nullptr && from_index != nullptr && eps != nullptr) {
            u = __ldg(sel_u01 + v);
            idx = __ldg(from_index + v);
            e0 = __ldg(eps + v * 2);
            e1 = __ldg(eps + v * 2 + 1);
        } else {
            const uint64_t g = offset + (uint64_t)v;
            const U4 r = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), kStreamMisalign, 0u, (uint32_t)seed,
                                       (uint32_t)(seed >> 32));
            u = u01(r.x);
            idx = 4 + min((int)(u01(r.y) * (float)span), span - 1);
            box_muller(r.z, r.w, e0, e1);
        }
        if (!(u < prob)) continue;                             // warp-uniform
        const float oef = fminf(fmaxf(__fadd_rn(__fmul_rn(e0, 0.15f), __ldg(oef_dbv + v * W)), 0.05f), 0.8f);
        const float dbv = fminf(fmaxf(__fadd_rn(__fmul_rn(e1, 0.05f), __ldg(oef_dbv + v * W + 1)), 0.002f), 0.3f);
        const float hct = HCT ? __ldg(oef_dbv + v * W + 2) : P.hct;
        const VoxelPhys vp = voxel_phys<HCT>(P, oef, dbv, hct);
        float I = 0.f, dI = 0.f;
        if (P.full_model) tissue_eval<false, PATH>(P, s, ss, qc, vp.dw, vp.dw_k, I, dI);
        const TauSignal ts = tau_signal<false>(P, vp, my_tau, my_b, I, dI);
        if (live && lane > idx) signal[v * nt + lane] = ts.S;
    }
}

// Paired variant of the scheduled path (n_tau <= 16, fixed Hct): a warp takes TWO voxels per iteration.  The two
// quadratures still run one after the other on all 32 lanes, but everything that only needs n_tau lanes -- loads,
// per-voxel physics, exp / blood / mixing epilogue, gradient reduction, stores -- is done once for both, voxel 0 on
// lanes 0-15 and voxel 1 on lanes 16-31 (about 120 fewer warp instructions per voxel).
template <bool BWD>
__global__ void __launch_bounds__(kThreads, QB_FWD_MIN_BLOCKS) k_forward_pair(const __grid_constant__ QboldParams P,
                                                                              const float* __restrict__ oef_dbv,
                                                                              const float* __restrict__ g_signal,
                                                                              float* __restrict__ signal,
                                                                              float* __restrict__ g_oef_dbv, int64_t n,
                                                                              unsigned long long* __restrict__ work) {
    __shared__ SchedSmem ss;
    load_sched(P, ss);
    __syncthreads();
    const int lane = threadIdx.x & 31, half = lane >> 4, t = lane & 15;
    const int nt = P.n_tau;
    const bool live = t < nt;
    const int my_col = live ? P.col_of_tau[t] : -1;
    const float my_tau = live ? P.tau[t] : 0.f;
    const float my_b = live ? P.blood_b[t] : 0.f;
    const QuadCtx qc = make_quad_ctx<kSched>(P, ss, lane, my_col, my_tau);
    const int64_t npairs = (n + 1) >> 1;

    // pairs come from a device work counter (next index fetched before the current pair is processed): warps that
    // run ahead simply take more pairs, which evens out SM-to-SM rate differences and the tail of the launch
    for (int64_t pr = next_unit(work, lane), nxt; pr < npairs; pr = nxt) {
        nxt = next_unit(work, lane);
        const int64_t v = pr * 2 + half;
        const bool valid = v < n;
        float2 x = make_float2(0.f, 0.f);
        if (valid) x = __ldg(reinterpret_cast<const float2*>(oef_dbv) + v);
        float gs = 1.0f;
        if (BWD && g_signal != nullptr && live && valid) gs = __ldg(g_signal + v * nt + t);
        const VoxelPhys vp = voxel_phys<false>(P, x.x, x.y, P.hct);
        const float A_mine = qc.tau_ref15 * vp.dw;
        float I = 0.f, Dm = 0.f;
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            const float A = __shfl_sync(kFull, A_mine, h << 4);
            float vi, vd;
            tissue_sched<BWD>(qc.nph, qc.sa, A, lane, qc.ph_lo, qc.ph_hi, my_col, vi, vd);
            if (half == h) {
                I = vi;
                Dm = vd;
            }
        }
        const float a_t = 1.5f * (fabsf(my_tau) * vp.dw);
        float dI = 0.f;
        if (BWD) dI = (Dm * qc.tau_ref15 + qc.node0_d * a_t) * vp.dw_k;
        if (my_col >= 0) I += node0_value(P, a_t);
        const TauSignal ts = tau_signal<BWD>(P, vp, my_tau, my_b, I, dI);
        if (live && valid && signal != nullptr) signal[v * nt + t] = ts.S;
        if (BWD) {
            float go = live ? gs * ts.dS_doef : 0.f;
            float gd = live ? gs * ts.dS_ddbv : 0.f;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                go += __shfl_xor_sync(kFull, go, o);
                gd += __shfl_xor_sync(kFull, gd, o);
            }
            if (t == 0 && valid) *reinterpret_cast<float2*>(g_oef_dbv + v * 2) = make_float2(go, gd);
        }
    }
}


// Log-linear branch (full_model = False, reference signals.py:194-207): ~30 FLOP per 52-104 B, i.e. HBM-bound.
// One thread per voxel; the [256 x n_tau] signal / upstream-gradient tiles go through shared memory so that
// every global access is a coalesced 16-byte vector (a CTA's rows are one contiguous 256*n_tau*4-byte span).
template <bool BWD>
__global__ void __launch_bounds__(kThreads) k_loglinear(const __grid_constant__ QboldParams P,
                                                        const float2* __restrict__ oef_dbv,
                                                        const float* __restrict__ g_signal, float* __restrict__ signal,
                                                        float2* __restrict__ g_oef_dbv, int64_t n) {
#ifdef QB_HOST_EMU
    float4* tile4 = reinterpret_cast<float4*>(qb_emu::dynamic_smem());   // tests/host_emu: the launch's dynamic window
#else
    extern __shared__ float4 tile4[];
#endif
    float* tile = reinterpret_cast<float*>(tile4);
    const int nt = P.n_tau;
    const int64_t v0 = (int64_t)blockIdx.x * kThreads;
    const int rows = (int)((n - v0 < kThreads) ? (n - v0) : kThreads);
    const int64_t base = v0 * nt;
    const int count = rows * nt;
    const bool vec = ((count & 3) == 0);                       // only the last CTA can be ragged
    const bool vec_g = vec && ((reinterpret_cast<uintptr_t>(g_signal + base) & 15) == 0);
    const bool vec_s = vec && ((reinterpret_cast<uintptr_t>(signal + base) & 15) == 0);
    if (BWD && g_signal != nullptr) {
        if (vec_g) {
            const float4* src = reinterpret_cast<const float4*>(g_signal + base);
            for (int i = threadIdx.x; i < count / 4; i += kThreads) tile4[i] = __ldg(src + i);
        } else {
            for (int i = threadIdx.x; i < count; i += kThreads) tile[i] = __ldg(g_signal + base + i);
        }
        __syncthreads();
    }
    const int r = threadIdx.x;
    float go = 0.f, gd = 0.f;
    if (r < rows) {
        const float2 x = __ldg(oef_dbv + v0 + r);
        const VoxelPhys vp = voxel_phys<false>(P, x.x, x.y, P.hct);
        for (int t = 0; t < nt; ++t) {
            const TauSignal ts = tau_signal<BWD>(P, vp, P.tau[t], P.blood_b[t], 0.f, 0.f);
            if (BWD) {
                const float gs = (g_signal != nullptr) ? tile[r * nt + t] : 1.0f;
                go = fmaf(gs, ts.dS_doef, go);
                gd = fmaf(gs, ts.dS_ddbv, gd);
            }
            tile[r * nt + t] = ts.S;                           // row stride n_tau (odd for 11): conflict-free
        }
        if (BWD) g_oef_dbv[v0 + r] = make_float2(go, gd);
    }
    if (signal == nullptr) return;
    __syncthreads();
    if (vec_s) {
        float4* dst = reinterpret_cast<float4*>(signal + base);
        for (int i = threadIdx.x; i < count / 4; i += kThreads) dst[i] = tile4[i];
    } else {
        for (int i = threadIdx.x; i < count; i += kThreads) signal[base + i] = tile[i];
    }
}

}  // namespace qb
