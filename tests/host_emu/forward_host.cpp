// TEST INFRASTRUCTURE: the forward kernels of the hot path (qbold_vi_b200/csrc/forward_kernels.cuh with qbold_core.cuh,
// bessel.cuh, rng.cuh -- the very source libqbold.so is built from) compiled for the host and run in the SIMT emulator
// of tests/host_emu/shim/cuda_runtime.h.  The CPU suite feeds them the parameter block of the real qbold_params_init and
// compares signal and gradient with the oracle: kernel arithmetic AND warp-level orchestration (lane schedule, phase
// flushes through shared memory, shuffles, work counter, the two-voxel pairing) are checked without a GPU.
//
//   g++ -O1 -std=c++17 -pthread -DQB_HOST_EMU -I tests/host_emu/shim -I qbold_vi_b200/csrc -shared -fPIC ...
#include <cuda_runtime.h>      // the shim

#include "forward_kernels.cuh"

namespace {
unsigned long long g_work;
}

// k_forward_pair<BWD> (the headline kernel): grid x block threads, block a multiple of 32.
QB_EMU_API void qb_emu_forward_pair(const QboldParams* P, const float* oef_dbv, const float* g_signal, float* signal,
                                    float* g_oef_dbv, int64_t n, int bwd, int grid, int block) {
    g_work = 0;
    const QboldParams params = *P;
    if (bwd)
        qb_emu::launch(grid, block, [&]() { qb::k_forward_pair<true>(params, oef_dbv, g_signal, signal, g_oef_dbv, n, &g_work); });
    else
        qb_emu::launch(grid, block, [&]() { qb::k_forward_pair<false>(params, oef_dbv, g_signal, signal, g_oef_dbv, n, &g_work); });
}

// k_forward<BWD, HCT, PATH>: one warp per voxel; path 0 = static lane schedule, 1 = column groups (<= 8 columns),
// 2 = column groups (> 8 columns, the 24-tau grid)
QB_EMU_API int qb_emu_forward(const QboldParams* P, const float* oef_dbv, const float* g_signal, float* signal,
                              float* g_oef_dbv, int64_t n, int bwd, int hct, int path, int grid, int block) {
    g_work = 0;
    const QboldParams params = *P;
#define QB_CASE(B, H, PA)                                                                                             \
    if (bwd == B && hct == H && path == PA) {                                                                         \
        qb_emu::launch(grid, block, [&]() {                                                                           \
            qb::k_forward<B != 0, H != 0, PA>(params, oef_dbv, g_signal, signal, g_oef_dbv, n, &g_work);              \
        });                                                                                                           \
        return 0;                                                                                                     \
    }
    QB_CASE(1, 0, 0) QB_CASE(0, 0, 0) QB_CASE(1, 1, 0) QB_CASE(1, 0, 1) QB_CASE(0, 0, 1) QB_CASE(1, 0, 2) QB_CASE(1, 1, 1)
#undef QB_CASE
    return -1;
}

// k_misalign<HCT, kSched>: overwrites the late images of the selected voxels in `signal` (recorded draws or Philox)
QB_EMU_API void qb_emu_misalign(const QboldParams* P, const float* oef_dbv, int64_t n, float prob, const float* sel_u01,
                                const int32_t* from_index, const float* eps, uint64_t seed, uint64_t offset, float* signal,
                                int hct, int grid, int block) {
    g_work = 0;
    const QboldParams params = *P;
    if (hct)
        qb_emu::launch(grid, block, [&]() {
            qb::k_misalign<true, qb::kSched>(params, oef_dbv, n, prob, sel_u01, from_index, eps, seed, offset, signal, &g_work);
        });
    else
        qb_emu::launch(grid, block, [&]() {
            qb::k_misalign<false, qb::kSched>(params, oef_dbv, n, prob, sel_u01, from_index, eps, seed, offset, signal, &g_work);
        });
}

// k_loglinear<BWD> (full_model = False, signals.py:194-207): one CTA of 256 rows per block, [256 x n_tau] tile in
// dynamic shared memory -- launch arithmetic of launch_loglinear (forward.cu)
QB_EMU_API void qb_emu_loglinear(const QboldParams* P, const float* oef_dbv, const float* g_signal, float* signal,
                                 float* g_oef_dbv, int64_t n, int bwd) {
    const QboldParams params = *P;
    const size_t smem = sizeof(float) * qb::kThreads * params.n_tau;
    const int grid = (int)((n + qb::kThreads - 1) / qb::kThreads);
    if (bwd)
        qb_emu::launch(grid, qb::kThreads, [&]() {
            qb::k_loglinear<true>(params, reinterpret_cast<const float2*>(oef_dbv), g_signal, signal,
                                  reinterpret_cast<float2*>(g_oef_dbv), n);
        }, 1, smem);
    else
        qb_emu::launch(grid, qb::kThreads, [&]() {
            qb::k_loglinear<false>(params, reinterpret_cast<const float2*>(oef_dbv), g_signal, signal,
                                   reinterpret_cast<float2*>(g_oef_dbv), n);
        }, 1, smem);
}
