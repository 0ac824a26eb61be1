// Host-buffer entry point: the same forward + VJP as qbold_forward_backward, fed from and
// returning to HOST memory.  The voxel range is cut into chunks that cycle over four
// stream slots so the H2D copy of chunk c+1, the kernel of chunk c and the D2H copy of
// chunk c-1 overlap (PCIe Gen5 full duplex).  Pinned host buffers give true asynchrony;
// pageable ones still work (the runtime stages them).
#include <mutex>

#include "launch.h"

namespace qb {

constexpr int kSlots = 6;
constexpr int64_t kChunk = 1 << 17;   // voxels per chunk: 1 MB in, 11.5 MB g+S, 1 MB grad: the pipeline fill + drain (one chunk each of H2D,
                                      // kernel, D2H) is 2 % of a 16 M-voxel call; with 512 k-voxel chunks it was 8 %

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

// What the box's host <-> device link sustains with the SAME copy pattern as qbold_forward_backward_host (pinned
// buffers, kChunk-voxel pieces cycling over the slot streams, both directions at once), without any kernel: the
// ceiling the end-to-end figure is judged against.  h_src / h_dst: host buffers of `bytes` each (pinned for a
// meaningful number); gbps[0] = H2D, gbps[1] = D2H, each direction's bytes / wall time of the concurrent run.
extern "C" int qbold_host_copy_ceiling(const void* h_src, void* h_dst, int64_t bytes, int32_t reps, double* gbps) {
    if (!h_src || !h_dst || !gbps || bytes <= 0 || reps < 1)
        return fail(QBOLD_EINVAL, "qbold_host_copy_ceiling: bad argument");
    std::lock_guard<std::mutex> lock(g_pipe_mu);
    int dev = 0;
    int rc = cuda_check(cudaGetDevice(&dev), "cudaGetDevice");
    if (rc) return rc;
    const int64_t piece = (int64_t)sizeof(float) * 13 * kChunk;          // one chunk's bytes per direction (8 + 44 B/voxel)
    void* d_in[kSlots] = {};
    void* d_out[kSlots] = {};
    cudaStream_t st[kSlots] = {};
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    for (int i = 0; i < kSlots && !rc; ++i) {
        rc = cuda_check(cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking), "cudaStreamCreate");
        if (!rc) rc = cuda_check(cudaMalloc(&d_in[i], piece), "cudaMalloc");
        if (!rc) rc = cuda_check(cudaMalloc(&d_out[i], piece), "cudaMalloc");
        if (!rc) rc = cuda_check(cudaMemsetAsync(d_out[i], 0, piece, st[i]), "cudaMemset");
    }
    if (!rc) rc = cuda_check(cudaEventCreate(&e0), "cudaEventCreate");
    if (!rc) rc = cuda_check(cudaEventCreate(&e1), "cudaEventCreate");
    float best_ms = 0.f;
    for (int r = 0; r <= reps && !rc; ++r) {                              // first pass is the warm-up
        rc = cuda_check(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
        if (!rc) rc = cuda_check(cudaEventRecord(e0, st[0]), "cudaEventRecord");
        for (int i = 1; i < kSlots && !rc; ++i) rc = cuda_check(cudaStreamWaitEvent(st[i], e0, 0), "cudaStreamWaitEvent");
        int64_t c = 0;
        for (int64_t off = 0; off < bytes && !rc; off += piece, ++c) {
            const int64_t m = bytes - off < piece ? bytes - off : piece;
            const int i = (int)(c % kSlots);
            rc = cuda_check(cudaMemcpyAsync(d_in[i], (const char*)h_src + off, m, cudaMemcpyHostToDevice, st[i]), "H2D");
            if (!rc) rc = cuda_check(cudaMemcpyAsync((char*)h_dst + off, d_out[i], m, cudaMemcpyDeviceToHost, st[i]), "D2H");
        }
        for (int i = 1; i < kSlots && !rc; ++i) {
            cudaEvent_t done;
            rc = cuda_check(cudaEventCreateWithFlags(&done, cudaEventDisableTiming), "cudaEventCreate");
            if (!rc) rc = cuda_check(cudaEventRecord(done, st[i]), "cudaEventRecord");
            if (!rc) rc = cuda_check(cudaStreamWaitEvent(st[0], done, 0), "cudaStreamWaitEvent");
            cudaEventDestroy(done);
        }
        if (!rc) rc = cuda_check(cudaEventRecord(e1, st[0]), "cudaEventRecord");
        if (!rc) rc = cuda_check(cudaEventSynchronize(e1), "cudaEventSynchronize");
        float ms = 0.f;
        if (!rc) rc = cuda_check(cudaEventElapsedTime(&ms, e0, e1), "cudaEventElapsedTime");
        if (r > 0 && (best_ms == 0.f || ms < best_ms)) best_ms = ms;
    }
    for (int i = 0; i < kSlots; ++i) {
        if (st[i]) cudaStreamDestroy(st[i]);
        cudaFree(d_in[i]);
        cudaFree(d_out[i]);
    }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (rc) return rc;
    gbps[0] = gbps[1] = (double)bytes / (best_ms * 1e-3) / 1e9;
    return QBOLD_OK;
}
