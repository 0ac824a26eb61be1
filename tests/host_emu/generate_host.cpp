// TEST INFRASTRUCTURE: the streaming synthetic-data kernels (qbold_vi_b200/csrc/generate_kernels.cuh; the source
// libqbold.so is built from) compiled for the host and run in the SIMT emulator.  The two functions below restate the
// few lines of launch logic of qbold_generate / qbold_add_noise_chunked (generate.cu) around the kernels.
#include <cuda_runtime.h>      // the shim

#include <vector>

#include "generate_kernels.cuh"

namespace {
unsigned long long g_work;
}

// pair != 0: k_generate_pair (n_tau <= 16, full model, scheduled path); else k_generate<path>
QB_EMU_API int qb_emu_generate(const QboldParams* P, const float* oefs, int64_t n_oef, const float* dbvs, int64_t n_dbv,
                               const int64_t* perm, uint64_t seed, int64_t first, int64_t count, float* x, float* y3,
                               int pair, int path, int grid) {
    g_work = 0;
    const QboldParams params = *P;
    const uint64_t total = (uint64_t)n_oef * (uint64_t)n_dbv;
    int bits = 1;
    while (bits < 64 && (1ull << bits) < total) ++bits;
    const int half_bits = (bits + 1) / 2;
    const int block = qb::kThreads;
    if (pair) {
        qb_emu::launch(grid, block, [&]() {
            qb::k_generate_pair(params, oefs, n_oef, dbvs, n_dbv, perm, seed, half_bits, first, count, x, y3, &g_work);
        });
        return 0;
    }
#define QB_CASE(PA)                                                                                                   \
    if (path == PA) {                                                                                                 \
        qb_emu::launch(grid, block, [&]() {                                                                           \
            qb::k_generate<PA>(params, oefs, n_oef, dbvs, n_dbv, perm, seed, half_bits, first, count, x, y3, &g_work); \
        });                                                                                                           \
        return 0;                                                                                                     \
    }
    QB_CASE(0) QB_CASE(1) QB_CASE(2)
#undef QB_CASE
    return -1;
}

// noise of create_synthetic_dataset's chunk loop (signals.py:116-128, 282-285): per-chunk column sums, then the noise pass
QB_EMU_API void qb_emu_add_noise_chunked(const QboldParams* P, float* signal, int64_t chunk_rows, int n_chunks,
                                         const float* snr_u01, const float* eps, uint64_t seed, uint64_t offset, int gx) {
    const QboldParams params = *P;
    const int64_t n = chunk_rows * n_chunks;
    std::vector<double> scratch((size_t)32 * n_chunks, 0.0);
    double* sums = scratch.data();
    const int nt = params.n_tau;
    qb_emu::launch(gx, qb::kThreads, [&]() { qb::k_column_sum_chunked(signal, chunk_rows, nt, sums); }, n_chunks);
    qb_emu::launch((int)((n + qb::kThreads - 1) / qb::kThreads), qb::kThreads, [&]() {
        qb::k_add_noise_chunked(params, signal, n, chunk_rows, sums, snr_u01, eps, seed, offset);
    });
}

// noise of one SignalGenerationLayer.call (signals.py:116-128): column means of the whole batch (qbold_column_mean:
// Kahan partial sums per warp, double totals), then one noise pass (qbold_add_noise).  mean_out receives the means.
QB_EMU_API void qb_emu_add_noise(const QboldParams* P, float* signal, int64_t n, const float* snr_u01, const float* eps,
                                 uint64_t seed, uint64_t offset, float* mean_out, int grid) {
    const QboldParams params = *P;
    const int nt = params.n_tau;
    std::vector<double> scratch(32, 0.0);
    double* sums = scratch.data();
    qb_emu::launch(grid, qb::kThreads, [&]() { qb::k_column_sum(signal, n, nt, sums); });
    qb_emu::launch(1, 32, [&]() { qb::k_finish_mean(sums, n, nt, mean_out); });
    qb_emu::launch((int)((n + qb::kThreads - 1) / qb::kThreads), qb::kThreads,
                   [&]() { qb::k_add_noise(params, signal, n, mean_out, snr_u01, eps, seed, offset); });
}
