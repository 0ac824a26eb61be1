// K3: streaming synthetic-data path.  Replaces the body of create_synthetic_dataset
// (reference signals.py:270-299) and the noise model of SignalGenerationLayer.call
// (signals.py:116-128) with three device passes per chunk:
//   k_generate     meshgrid('ij') + shuffle + forward model + labels (OEF, DBV, R2')
//   k_column_sum   the batch statistic mean_over_chunk(signal) of signals.py:126
//   k_add_noise    snr ~ U(50,120) * norm_snr, signal += N(0,1) * mean/snr   (HBM-bound pass)
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

using namespace qb;

extern "C" int qbold_add_noise_chunked(const QboldParams* p, float* signal, int64_t chunk_rows, int32_t n_chunks,
                                       const float* snr_u01, const float* eps, uint64_t seed, uint64_t offset,
                                       double* scratch, void* stream) {
    if (!p || p->abi_version != QBOLD_ABI_VERSION)
        return fail(QBOLD_EINVAL, "qbold_add_noise_chunked: bad params block");
    if (p->norm_snr[0] == 0.0f)
        return fail(QBOLD_EUNSUPPORTED,
                    "norm_snr is only defined for 11 or 24 taus (reference signals.py:117-121), got %d", p->n_tau);
    if ((snr_u01 == nullptr) != (eps == nullptr))
        return fail(QBOLD_EINVAL, "qbold_add_noise_chunked: pass both snr_u01 and eps, or neither");
    if (chunk_rows < 0 || n_chunks < 0 || n_chunks > 65535)
        return fail(QBOLD_EINVAL, "qbold_add_noise_chunked: bad chunking");
    const int64_t n = chunk_rows * n_chunks;
    if (n == 0) return QBOLD_OK;
    if (!signal || !scratch) return fail(QBOLD_EINVAL, "qbold_add_noise_chunked: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = cuda_check(cudaMemsetAsync(scratch, 0, sizeof(double) * 32 * n_chunks, st), "cudaMemsetAsync");
    if (rc) return rc;
    int64_t gx = (chunk_rows + 8 * 64 - 1) / (8 * 64);
    const int64_t cap = ((int64_t)sm_count() * 8 + n_chunks - 1) / n_chunks;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    k_column_sum_chunked<<<dim3((unsigned)gx, (unsigned)n_chunks), kThreads, 0, st>>>(signal, chunk_rows, p->n_tau,
                                                                                    scratch);
    rc = after_launch("k_column_sum_chunked");
    if (rc) return rc;
    k_add_noise_chunked<<<(unsigned)((n + kThreads - 1) / kThreads), kThreads, 0, st>>>(*p, signal, n, chunk_rows, scratch,
                                                                                       snr_u01, eps, seed, offset);
    return after_launch("k_add_noise_chunked");
}

extern "C" int qbold_generate(const QboldParams* p, const float* oefs, int64_t n_oef, const float* dbvs,
                              int64_t n_dbv, const int64_t* perm, uint64_t seed, int64_t first, int64_t count,
                              float* x, float* y3, void* stream) {
    if (!p || p->abi_version != QBOLD_ABI_VERSION) return fail(QBOLD_EINVAL, "qbold_generate: bad params block");
    if (!oefs || !dbvs || n_oef < 1 || n_dbv < 1) return fail(QBOLD_EINVAL, "qbold_generate: empty marginals");
    if (first < 0 || count < 0 || first + count > n_oef * n_dbv)
        return fail(QBOLD_EINVAL, "qbold_generate: rows [%lld,%lld) outside the %lld x %lld meshgrid",
                    (long long)first, (long long)(first + count), (long long)n_oef, (long long)n_dbv);
    if (count == 0) return QBOLD_OK;
    if (!x && !y3) return fail(QBOLD_EINVAL, "qbold_generate: no output requested");
    const uint64_t total = (uint64_t)n_oef * (uint64_t)n_dbv;
    int bits = 1;
    while (bits < 64 && (1ull << bits) < total) ++bits;
    const int half_bits = (bits + 1) / 2;
    const int path = p->sched_phases > 0 ? kSched : (p->n_cols > kColGroup ? kColsMulti : kCols);
    static int bps_cache[3] = {0, 0, 0};
    int& bps = bps_cache[path];
    if (bps == 0) {
        cudaError_t e;
        if (path == kSched) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_generate<kSched>, kThreads, 0);
        else if (path == kCols) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_generate<kCols>, kThreads, 0);
        else e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_generate<kColsMulti>, kThreads, 0);
        if (e != cudaSuccess || bps < 1) bps = 1;
    }
    int64_t grid = (int64_t)sm_count() * bps;
    const int64_t want = (count + 7) / 8;
    if (want < grid) grid = want;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* gwork = next_work_counter(st);
    if (!gwork) return fail(QBOLD_ECUDA, "qbold_generate: work counter unavailable");
    if (path == kSched && p->full_model && p->n_tau <= 16) {
        static int bps_pair = 0;
        if (bps_pair == 0 && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps_pair, k_generate_pair, kThreads, 0) !=
                                  cudaSuccess || bps_pair < 1))
            bps_pair = 1;
        int64_t gp = (int64_t)sm_count() * bps_pair;
        const int64_t wantp = ((count + 1) / 2 + 7) / 8;
        if (wantp < gp) gp = wantp;
        unsigned long long* work = gwork;
        k_generate_pair<<<(unsigned)gp, kThreads, 0, st>>>(*p, oefs, n_oef, dbvs, n_dbv, perm, seed, half_bits, first,
                                                           count, x, y3, work);
    } else if (path == kSched)
        k_generate<kSched><<<(unsigned)grid, kThreads, 0, st>>>(*p, oefs, n_oef, dbvs, n_dbv, perm, seed, half_bits,
                                                                first, count, x, y3, gwork);
    else if (path == kCols)
        k_generate<kCols><<<(unsigned)grid, kThreads, 0, st>>>(*p, oefs, n_oef, dbvs, n_dbv, perm, seed, half_bits,
                                                               first, count, x, y3, gwork);
    else
        k_generate<kColsMulti><<<(unsigned)grid, kThreads, 0, st>>>(*p, oefs, n_oef, dbvs, n_dbv, perm, seed,
                                                                    half_bits, first, count, x, y3, gwork);
    return after_launch("k_generate");
}

extern "C" int qbold_column_mean(const float* signal, int64_t n, int32_t n_tau, float* mean, double* scratch,
                                 void* stream) {
    if (!signal || !mean || !scratch || n < 1 || n_tau < 1 || n_tau > 32)
        return fail(QBOLD_EINVAL, "qbold_column_mean: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = cuda_check(cudaMemsetAsync(scratch, 0, sizeof(double) * n_tau, st), "cudaMemsetAsync");
    if (rc) return rc;
    int64_t grid = (int64_t)sm_count() * 4;
    const int64_t want = (n + 7) / 8;
    if (want < grid) grid = want;
    k_column_sum<<<(unsigned)grid, kThreads, 0, st>>>(signal, n, n_tau, scratch);
    rc = after_launch("k_column_sum");
    if (rc) return rc;
    k_finish_mean<<<1, 32, 0, st>>>(scratch, n, n_tau, mean);
    return after_launch("k_finish_mean");
}

extern "C" int qbold_add_noise(const QboldParams* p, float* signal, int64_t n, const float* mean,
                               const float* snr_u01, const float* eps, uint64_t seed, uint64_t offset,
                               void* stream) {
    if (!p || p->abi_version != QBOLD_ABI_VERSION) return fail(QBOLD_EINVAL, "qbold_add_noise: bad params block");
    if (p->norm_snr[0] == 0.0f)
        return fail(QBOLD_EUNSUPPORTED,
                    "norm_snr is only defined for 11 or 24 taus (reference signals.py:117-121), got %d", p->n_tau);
    if ((snr_u01 == nullptr) != (eps == nullptr))
        return fail(QBOLD_EINVAL, "qbold_add_noise: pass both snr_u01 and eps, or neither");
    if (n < 0 || (n > 0 && (!signal || !mean))) return fail(QBOLD_EINVAL, "qbold_add_noise: null pointer");
    if (n == 0) return QBOLD_OK;
    k_add_noise<<<(unsigned)((n + kThreads - 1) / kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        *p, signal, n, mean, snr_u01, eps, seed, offset);
    return after_launch("k_add_noise");
}
