// K3: streaming synthetic-data path.  Replaces the body of create_synthetic_dataset
// (reference signals.py:270-299) and the noise model of SignalGenerationLayer.call
// (signals.py:116-128) with three device passes per chunk:
//   k_generate     meshgrid('ij') + shuffle + forward model + labels (OEF, DBV, R2')
//   k_column_sum   the batch statistic mean_over_chunk(signal) of signals.py:126
//   k_add_noise    snr ~ U(50,120) * norm_snr, signal += N(0,1) * mean/snr   (HBM-bound pass)
//
// The kernels live in this header, apart from their launchers in generate.cu, so that tests/host_emu can compile the
// same kernel source for the host and run it in its SIMT emulator (CPU suite).
#pragma once
#include "qbold_core.cuh"
#include "launch.h"
#include "rng.cuh"

namespace qb {

// (the keyed Feistel bijection that stands in for tf.random.shuffle lives in rng.cuh: feistel_permute)

template <int PATH>
__global__ void __launch_bounds__(kThreads) k_generate(const __grid_constant__ QboldParams P,
                                                       const float* __restrict__ oefs, int64_t n_oef,
                                                       const float* __restrict__ dbvs, int64_t n_dbv,
                                                       const int64_t* __restrict__ perm, uint64_t seed,
                                                       int half_bits, int64_t first, int64_t count,
                                                       float* __restrict__ x, float* __restrict__ y3,
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
    const QuadCtx qc = make_quad_ctx<PATH>(P, ss, lane, my_col, my_tau);
    const uint64_t total = (uint64_t)n_oef * (uint64_t)n_dbv;

    for (int64_t v = next_unit(work, lane), nxt_unit; v < count; v = nxt_unit) {   // dynamic units, see next_unit
        nxt_unit = next_unit(work, lane);
        const uint64_t row = (uint64_t)(first + v);
        const uint64_t idx = perm ? (uint64_t)__ldg(perm + row) : feistel_permute(row, total, half_bits, seed);
        const float oef = __ldg(oefs + idx / (uint64_t)n_dbv);            // meshgrid(indexing='ij'), signals.py:270
        const float dbv = __ldg(dbvs + idx % (uint64_t)n_dbv);
        if (y3 != nullptr && lane < 3) {
            const float r2p = (P.dw_k * oef) * dbv;                        // signals.py:296
            y3[v * 3 + lane] = lane == 0 ? oef : (lane == 1 ? dbv : r2p);
        }
        if (x == nullptr) continue;
        const VoxelPhys vp = voxel_phys<false>(P, oef, dbv, P.hct);
        float I = 0.f, dI = 0.f;
        if (P.full_model) tissue_eval<false, PATH>(P, s, ss, qc, vp.dw, vp.dw_k, I, dI);
        const TauSignal ts = tau_signal<false>(P, vp, my_tau, my_b, I, dI);
        if (live) x[v * nt + lane] = ts.S;
    }
}

// Paired variant (scheduled path, n_tau <= 16): two voxels per warp iteration, voxel 0 on lanes 0-15 and voxel 1 on
// lanes 16-31 for everything but the two quadratures (see k_forward_pair).
// idx -> (row of oefs, row of dbvs) of the 'ij' meshgrid; 32-bit division when the grid has < 2^32 points.
__device__ __forceinline__ void mesh_index(uint64_t idx, uint64_t n_dbv, bool small, uint64_t& io, uint64_t& id) {
    if (small) {
        const unsigned q = (unsigned)idx / (unsigned)n_dbv;
        io = q;
        id = (unsigned)idx - q * (unsigned)n_dbv;
    } else {
        io = idx / n_dbv;
        id = idx % n_dbv;
    }
}

__global__ void __launch_bounds__(kThreads, 5) k_generate_pair(const __grid_constant__ QboldParams P,
                                                               const float* __restrict__ oefs, int64_t n_oef,
                                                               const float* __restrict__ dbvs, int64_t n_dbv,
                                                               const int64_t* __restrict__ perm, uint64_t seed,
                                                               int half_bits, int64_t first, int64_t count,
                                                               float* __restrict__ x, float* __restrict__ y3,
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
    const uint64_t total = (uint64_t)n_oef * (uint64_t)n_dbv;
    const int64_t npairs = (count + 1) >> 1;

    for (int64_t pr = next_unit(work, lane), nxt; pr < npairs; pr = nxt) {   // dynamic pairs, see k_forward_pair
        nxt = next_unit(work, lane);
        int64_t v = pr * 2 + half;
        const bool valid = v < count;
        if (!valid) v = count - 1;
        const uint64_t row = (uint64_t)(first + v);
        const uint64_t idx = perm ? (uint64_t)__ldg(perm + row) : feistel_permute(row, total, half_bits, seed);
        uint64_t io, id;
        mesh_index(idx, (uint64_t)n_dbv, total < (1ull << 32), io, id);
        const float oef = __ldg(oefs + io);                               // meshgrid(indexing='ij'), signals.py:270
        const float dbv = __ldg(dbvs + id);
        if (y3 != nullptr && t < 3 && valid) {
            const float r2p = (P.dw_k * oef) * dbv;                        // signals.py:296
            y3[v * 3 + t] = t == 0 ? oef : (t == 1 ? dbv : r2p);
        }
        if (x == nullptr) continue;
        const VoxelPhys vp = voxel_phys<false>(P, oef, dbv, P.hct);
        const float A_mine = qc.tau_ref15 * vp.dw;
        float I = 0.f;
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            const float A = __shfl_sync(kFull, A_mine, h << 4);
            float vi, vd;
            tissue_sched<false>(qc.nph, qc.sa, A, lane, qc.ph_lo, qc.ph_hi, my_col, vi, vd);
            if (half == h) I = vi;
        }
        if (my_col >= 0) I += node0_value(P, 1.5f * (fabsf(my_tau) * vp.dw));
        const TauSignal ts = tau_signal<false>(P, vp, my_tau, my_b, I, 0.f);
        if (live && valid) x[v * nt + t] = ts.S;
    }
}

// Column sums of a [n, nt] block; one warp reads one 4*nt-byte row per step.
__global__ void __launch_bounds__(kThreads) k_column_sum(const float* __restrict__ sig, int64_t n, int nt,
                                                         double* __restrict__ sums) {
    __shared__ float part[kThreads / 32][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + w;
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    float acc = 0.f, comp = 0.f;   // Kahan: rows per warp can reach 1e5+
    if (lane < nt) {
        for (int64_t r = warp; r < n; r += nwarps) {
            const float yv = __ldg(sig + r * nt + lane) - comp;
            const float t = acc + yv;
            comp = (t - acc) - yv;
            acc = t;
        }
    }
    part[w][lane] = acc;
    __syncthreads();
    if (w == 0 && lane < nt) {
        double tot = 0.0;
        for (int k = 0; k < kThreads / 32; ++k) tot += (double)part[k][lane];
        atomicAdd(sums + lane, tot);
    }
}

__global__ void k_finish_mean(const double* __restrict__ sums, int64_t n, int nt, float* __restrict__ mean) {
    const int t = threadIdx.x;
    if (t < nt) mean[t] = (float)(sums[t] / (double)n);
}

// signals.py:116-128, one thread per voxel (rows stay L1-resident across the tau loop).
__global__ void __launch_bounds__(kThreads) k_add_noise(const __grid_constant__ QboldParams P, float* __restrict__ sig,
                                                        int64_t n, const float* __restrict__ mean,
                                                        const float* __restrict__ snr_u01,
                                                        const float* __restrict__ eps, uint64_t seed, uint64_t offset) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const int nt = P.n_tau;
    float u;
    if (snr_u01) {
        u = snr_u01[v];
    } else {
        const U4 r = philox4x32_10((uint32_t)(offset + v), (uint32_t)((offset + v) >> 32), kStreamSnr, 0u,
                                   (uint32_t)seed, (uint32_t)(seed >> 32));
        u = u01(r.x);
    }
    const float snr0 = u * (120.0f - 50.0f) + 50.0f;                       // tf.random.uniform(.., 50, 120), :124
    for (int t = 0; t < nt; t += 2) {
        float n0, n1;
        if (eps) {
            n0 = eps[v * nt + t];
            n1 = (t + 1 < nt) ? eps[v * nt + t + 1] : 0.f;
        } else {
            normal_pair(seed, offset + (uint64_t)v, kStreamNoise + (uint32_t)(t >> 1), n0, n1);
        }
        const float sd0 = __ldg(mean + t) / (snr0 * P.norm_snr[t]);        // :124-126
        sig[v * nt + t] = sig[v * nt + t] + n0 * sd0;                      // :128
        if (t + 1 < nt) {
            const float sd1 = __ldg(mean + t + 1) / (snr0 * P.norm_snr[t + 1]);
            sig[v * nt + t + 1] = sig[v * nt + t + 1] + n1 * sd1;
        }
    }
}

// All chunks of create_synthetic_dataset's noise loop (signals.py:282-285: every chunk of S^2/10 rows is one forward
// call, so the noise std uses THAT chunk's column means) in two launches: blockIdx.y = chunk for the column sums,
// then one noise pass that looks its chunk's sums up.  Same per-row draws and arithmetic as k_add_noise.
__global__ void __launch_bounds__(kThreads) k_column_sum_chunked(const float* __restrict__ sig, int64_t chunk_rows, int nt,
                                                                 double* __restrict__ sums) {
    __shared__ float part[kThreads / 32][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const float* base = sig + (int64_t)blockIdx.y * chunk_rows * nt;
    const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + w;
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    float acc = 0.f, comp = 0.f;
    if (lane < nt) {
        for (int64_t r = warp; r < chunk_rows; r += nwarps) {
            const float yv = __ldg(base + r * nt + lane) - comp;
            const float t = acc + yv;
            comp = (t - acc) - yv;
            acc = t;
        }
    }
    part[w][lane] = acc;
    __syncthreads();
    if (w == 0 && lane < nt) {
        double tot = 0.0;
        for (int k = 0; k < kThreads / 32; ++k) tot += (double)part[k][lane];
        atomicAdd(sums + blockIdx.y * 32 + lane, tot);
    }
}

__global__ void __launch_bounds__(kThreads) k_add_noise_chunked(const __grid_constant__ QboldParams P,
                                                                float* __restrict__ sig, int64_t n, int64_t chunk_rows,
                                                                const double* __restrict__ sums,
                                                                const float* __restrict__ snr_u01,
                                                                const float* __restrict__ eps, uint64_t seed,
                                                                uint64_t offset) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const int nt = P.n_tau;
    const double* cs = sums + (v / chunk_rows) * 32;
    float u;
    if (snr_u01) {
        u = snr_u01[v];
    } else {
        const U4 r = philox4x32_10((uint32_t)(offset + v), (uint32_t)((offset + v) >> 32), kStreamSnr, 0u,
                                   (uint32_t)seed, (uint32_t)(seed >> 32));
        u = u01(r.x);
    }
    const float snr0 = u * (120.0f - 50.0f) + 50.0f;
    for (int t = 0; t < nt; t += 2) {
        float n0, n1;
        if (eps) {
            n0 = eps[v * nt + t];
            n1 = (t + 1 < nt) ? eps[v * nt + t + 1] : 0.f;
        } else {
            normal_pair(seed, offset + (uint64_t)v, kStreamNoise + (uint32_t)(t >> 1), n0, n1);
        }
        const float m0 = (float)(cs[t] / (double)chunk_rows);
        sig[v * nt + t] = sig[v * nt + t] + n0 * (m0 / (snr0 * P.norm_snr[t]));
        if (t + 1 < nt) {
            const float m1 = (float)(cs[t + 1] / (double)chunk_rows);
            sig[v * nt + t + 1] = sig[v * nt + t + 1] + n1 * (m1 / (snr0 * P.norm_snr[t + 1]));
        }
    }
}

}  // namespace qb
