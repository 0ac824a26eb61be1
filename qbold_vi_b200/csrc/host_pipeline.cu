// Host-buffer entry point: the same forward + VJP as qbold_forward_backward, fed from and
// returning to HOST memory.  The voxel range is cut into chunks that cycle over six
// stream slots so the H2D copy of chunk c+1, the kernel of chunk c and the D2H copy of
// chunk c-1 overlap (PCIe Gen5 full duplex).  Pinned host buffers give true asynchrony;
// pageable ones still work (the runtime stages them).
#include <cstdlib>
#include <mutex>

#include "launch.h"

namespace qb {

constexpr int kSlots = 6;
// Voxels per chunk.  Two effects pull against each other: the pipeline's fill + drain (one chunk each of H2D, kernel,
// D2H) and the fixed cost per chunk (a kernel launch whose persistent grid ramps up and drains, four copies).  Measured
// for 16.7 M voxels on a B200 (tools/e2e_chunk_sweep.py): 2^17: 22.7 ms, 2^18: 21.0 ms, 2^19: 21.0 ms, 2^20: 21.9 ms
// against 19-19.6 ms of pure copies.  QBOLD_HOST_CHUNK_LOG2 overrides the default for sweeps.
static int64_t chunk_voxels() {
    static int64_t v = 0;
    if (v == 0) {
        const char* e = getenv("QBOLD_HOST_CHUNK_LOG2");
        const int lg = e ? atoi(e) : 0;
        v = (lg >= 12 && lg <= 22) ? ((int64_t)1 << lg) : ((int64_t)1 << 18);
    }
    return v;
}
#define kChunk (chunk_voxels())

// Chunks cycle over kSlots streams (g_signal in, signal out: 44 bytes per voxel each way); the 8-byte-per-voxel arrays
// (OEF/DBV in, its gradient out) travel per SUPER-chunk of kSuper chunks on two streams of their own, so a chunk costs
// two large copies instead of four.
constexpr int kSuperSlots = 3;
static int super_chunks() {                      // chunks per super-chunk: ~1 M voxels (8 MB each way)
    const int64_t k = ((int64_t)1 << 20) / chunk_voxels();
    return k < 1 ? 1 : (int)k;
}
#define kSuper (super_chunks())

struct Slot {
    cudaStream_t stream = nullptr;
    float *g = nullptr, *sig = nullptr;
    cudaEvent_t kernel_done = nullptr;
};

struct SuperSlot {
    float *in = nullptr, *grad = nullptr;
    cudaEvent_t in_ready = nullptr, grad_out = nullptr;
};

struct Pipeline {
    int device = -1;
    int n_tau = 0;
    int64_t chunk = 0;
    Slot slot[kSlots];
    SuperSlot sup[kSuperSlots];
    cudaStream_t in_stream = nullptr, grad_stream = nullptr;
};

static Pipeline g_pipe;
static std::mutex g_pipe_mu;

static void release(Pipeline& p) {
    for (auto& s : p.slot) {
        if (s.stream) cudaStreamDestroy(s.stream);
        if (s.kernel_done) cudaEventDestroy(s.kernel_done);
        cudaFree(s.g);
        cudaFree(s.sig);
        s = Slot{};
    }
    for (auto& u : p.sup) {
        if (u.in_ready) cudaEventDestroy(u.in_ready);
        if (u.grad_out) cudaEventDestroy(u.grad_out);
        cudaFree(u.in);
        cudaFree(u.grad);
        u = SuperSlot{};
    }
    if (p.in_stream) cudaStreamDestroy(p.in_stream);
    if (p.grad_stream) cudaStreamDestroy(p.grad_stream);
    p.in_stream = p.grad_stream = nullptr;
    p.device = -1;
}

static int ensure(Pipeline& p, int device, int n_tau) {
    if (p.device == device && p.n_tau == n_tau && p.chunk == kChunk) return QBOLD_OK;
    if (p.device >= 0) release(p);
    int rc = cuda_check(cudaStreamCreateWithFlags(&p.in_stream, cudaStreamNonBlocking), "cudaStreamCreate");
    if (!rc) rc = cuda_check(cudaStreamCreateWithFlags(&p.grad_stream, cudaStreamNonBlocking), "cudaStreamCreate");
    for (auto& s : p.slot) {
        if (!rc) rc = cuda_check(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking), "cudaStreamCreate");
        if (!rc) rc = cuda_check(cudaEventCreateWithFlags(&s.kernel_done, cudaEventDisableTiming), "cudaEventCreate");
        if (!rc) rc = cuda_check(cudaMalloc(&s.g, sizeof(float) * n_tau * kChunk), "cudaMalloc");
        if (!rc) rc = cuda_check(cudaMalloc(&s.sig, sizeof(float) * n_tau * kChunk), "cudaMalloc");
    }
    for (auto& u : p.sup) {
        if (!rc) rc = cuda_check(cudaEventCreateWithFlags(&u.in_ready, cudaEventDisableTiming), "cudaEventCreate");
        if (!rc) rc = cuda_check(cudaEventCreateWithFlags(&u.grad_out, cudaEventDisableTiming), "cudaEventCreate");
        if (!rc) rc = cuda_check(cudaMalloc(&u.in, sizeof(float) * 2 * kChunk * kSuper), "cudaMalloc");
        if (!rc) rc = cuda_check(cudaMalloc(&u.grad, sizeof(float) * 2 * kChunk * kSuper), "cudaMalloc");
    }
    if (rc) {
        release(p);
        return rc;
    }
    p.device = device;
    p.n_tau = n_tau;
    p.chunk = kChunk;
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
    Pipeline& P = g_pipe;
    const int64_t chunk = kChunk, super = chunk * kSuper;
    int64_t c = 0, sc = 0;
    for (int64_t sfirst = 0; sfirst < n && !rc; sfirst += super, ++sc) {
        const int64_t sm = (n - sfirst < super) ? (n - sfirst) : super;
        SuperSlot& u = P.sup[sc % kSuperSlots];
        // OEF/DBV of the whole super-chunk; the slot's previous gradient must have left first (its kernels are done then)
        if (sc >= kSuperSlots) rc = cuda_check(cudaStreamWaitEvent(P.in_stream, u.grad_out, 0), "cudaStreamWaitEvent");
        if (!rc) rc = cuda_check(cudaMemcpyAsync(u.in, h_oef_dbv + sfirst * 2, sizeof(float) * 2 * sm, cudaMemcpyHostToDevice,
                                                 P.in_stream), "H2D oef_dbv");
        if (!rc) rc = cuda_check(cudaEventRecord(u.in_ready, P.in_stream), "cudaEventRecord");
        for (int64_t first = sfirst; first < sfirst + sm && !rc; first += chunk, ++c) {
            const int64_t m = (sfirst + sm - first < chunk) ? (sfirst + sm - first) : chunk;
            Slot& s = P.slot[c % kSlots];
            const int64_t off = first - sfirst;
            rc = cuda_check(cudaStreamWaitEvent(s.stream, u.in_ready, 0), "cudaStreamWaitEvent");
            if (!rc && h_g_signal)
                rc = cuda_check(cudaMemcpyAsync(s.g, h_g_signal + first * nt, sizeof(float) * nt * m,
                                                cudaMemcpyHostToDevice, s.stream), "H2D g_signal");
            if (!rc) rc = qbold_forward_backward(p, u.in + off * 2, h_g_signal ? s.g : nullptr, m, h_signal ? s.sig : nullptr,
                                                 u.grad + off * 2, s.stream);
            if (!rc) rc = cuda_check(cudaEventRecord(s.kernel_done, s.stream), "cudaEventRecord");
            if (!rc) rc = cuda_check(cudaStreamWaitEvent(P.grad_stream, s.kernel_done, 0), "cudaStreamWaitEvent");
            if (!rc && h_signal)
                rc = cuda_check(cudaMemcpyAsync(h_signal + first * nt, s.sig, sizeof(float) * nt * m,
                                                cudaMemcpyDeviceToHost, s.stream), "D2H signal");
        }
        // the super-chunk's gradient once all its kernels have run
        if (!rc) rc = cuda_check(cudaMemcpyAsync(h_g_oef_dbv + sfirst * 2, u.grad, sizeof(float) * 2 * sm,
                                                 cudaMemcpyDeviceToHost, P.grad_stream), "D2H grad");
        if (!rc) rc = cuda_check(cudaEventRecord(u.grad_out, P.grad_stream), "cudaEventRecord");
    }
    for (auto& s : P.slot) {
        const int rc2 = cuda_check(cudaStreamSynchronize(s.stream), "cudaStreamSynchronize");
        if (!rc) rc = rc2;
    }
    const int rc3 = cuda_check(cudaStreamSynchronize(P.grad_stream), "cudaStreamSynchronize");
    if (!rc) rc = rc3;
    const int rc4 = cuda_check(cudaStreamSynchronize(P.in_stream), "cudaStreamSynchronize");
    if (!rc) rc = rc4;
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
