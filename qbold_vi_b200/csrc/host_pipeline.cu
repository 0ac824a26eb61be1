// Host-buffer entry point: the same forward + VJP as qbold_forward_backward, fed from and
// returning to HOST memory.  The voxel range is cut into chunks that cycle over four
// stream slots so the H2D copy of chunk c+1, the kernel of chunk c and the D2H copy of
// chunk c-1 overlap (PCIe Gen5 full duplex).  Pinned host buffers give true asynchrony;
// pageable ones still work (the runtime stages them).
#include <mutex>

#include "launch.h"

namespace qb {

constexpr int kSlots = 4;
constexpr int64_t kChunk = 1 << 19;   // voxels per chunk: 4 MB in, 46 MB g+S, 4 MB grad (short pipeline fill/drain)

struct Slot {
    cudaStream_t stream = nullptr;
    float *in = nullptr, *g = nullptr, *sig = nullptr, *grad = nullptr;
};

struct Pipeline {
    int device = -1;
    int n_tau = 0;
    Slot slot[kSlots];
};

static Pipeline g_pipe;
static std::mutex g_pipe_mu;

static void release(Pipeline& p) {
    for (auto& s : p.slot) {
        if (s.stream) cudaStreamDestroy(s.stream);
        cudaFree(s.in);
        cudaFree(s.g);
        cudaFree(s.sig);
        cudaFree(s.grad);
        s = Slot{};
    }
    p.device = -1;
}

static int ensure(Pipeline& p, int device, int n_tau) {
    if (p.device == device && p.n_tau == n_tau) return QBOLD_OK;
    if (p.device >= 0) release(p);
    for (auto& s : p.slot) {
        int rc = cuda_check(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking), "cudaStreamCreate");
        if (!rc) rc = cuda_check(cudaMalloc(&s.in, sizeof(float) * 2 * kChunk), "cudaMalloc");
        if (!rc) rc = cuda_check(cudaMalloc(&s.g, sizeof(float) * n_tau * kChunk), "cudaMalloc");
        if (!rc) rc = cuda_check(cudaMalloc(&s.sig, sizeof(float) * n_tau * kChunk), "cudaMalloc");
        if (!rc) rc = cuda_check(cudaMalloc(&s.grad, sizeof(float) * 2 * kChunk), "cudaMalloc");
        if (rc) {
            release(p);
            return rc;
        }
    }
    p.device = device;
    p.n_tau = n_tau;
    return QBOLD_OK;
}

}  // namespace qb

using namespace qb;

extern "C" int qbold_forward_backward_host(const QboldParams* p, const float* h_oef_dbv, const float* h_g_signal,
                                           int64_t n, float* h_signal, float* h_g_oef_dbv) {
    if (!p || p->abi_version != QBOLD_ABI_VERSION)
        return fail(QBOLD_EINVAL, "qbold_forward_backward_host: bad params block");
    if (n < 0 || (n > 0 && (!h_oef_dbv || !h_g_oef_dbv)))
        return fail(QBOLD_EINVAL, "qbold_forward_backward_host: null pointer");
    if (n == 0) return QBOLD_OK;
    std::lock_guard<std::mutex> lock(g_pipe_mu);
    int dev = 0;
    int rc = cuda_check(cudaGetDevice(&dev), "cudaGetDevice");
    if (rc) return rc;
    const int nt = p->n_tau;
    rc = ensure(g_pipe, dev, nt);
    if (rc) return rc;
    int64_t c = 0;
    for (int64_t first = 0; first < n && !rc; first += kChunk, ++c) {
        const int64_t m = (n - first < kChunk) ? (n - first) : kChunk;
        Slot& s = g_pipe.slot[c % kSlots];
        rc = cuda_check(cudaMemcpyAsync(s.in, h_oef_dbv + first * 2, sizeof(float) * 2 * m, cudaMemcpyHostToDevice,
                                        s.stream), "H2D oef_dbv");
        if (!rc && h_g_signal)
            rc = cuda_check(cudaMemcpyAsync(s.g, h_g_signal + first * nt, sizeof(float) * nt * m,
                                            cudaMemcpyHostToDevice, s.stream), "H2D g_signal");
        if (!rc) rc = qbold_forward_backward(p, s.in, h_g_signal ? s.g : nullptr, m, h_signal ? s.sig : nullptr,
                                             s.grad, s.stream);
        if (!rc && h_signal)
            rc = cuda_check(cudaMemcpyAsync(h_signal + first * nt, s.sig, sizeof(float) * nt * m,
                                            cudaMemcpyDeviceToHost, s.stream), "D2H signal");
        if (!rc)
            rc = cuda_check(cudaMemcpyAsync(h_g_oef_dbv + first * 2, s.grad, sizeof(float) * 2 * m,
                                            cudaMemcpyDeviceToHost, s.stream), "D2H grad");
    }
    for (auto& s : g_pipe.slot) {
        const int rc2 = cuda_check(cudaStreamSynchronize(s.stream), "cudaStreamSynchronize");
        if (!rc) rc = rc2;
    }
    return rc;
}
