// K3: streaming synthetic-data path -- launchers and C entry points (kernels: generate_kernels.cuh).
#include "generate_kernels.cuh"

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
