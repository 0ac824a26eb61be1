// Micro-benchmark: cycles per tcgen05.mma kind::tf32 (SS operands, K = 8) for several M x N shapes and operand majors,
// one CTA per SM, back-to-back issue from one thread.  Answers two design questions of the encoder kernels: what an
// M = 64 instruction costs against M = 128 (is a half-empty M = 128 better than M = 64?), and where the shared-memory
// operand bandwidth, not the tensor pipe, bounds a small-N instruction.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/micro/umma_rate tools/micro/umma_rate.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t mk_desc(unsigned addr, unsigned lbo, unsigned sbo, unsigned layout) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)layout << 61);
}

__global__ void __launch_bounds__(128) k_rate(int M, int N, int mn_major, int iters, int a_shift, int a_lbo, int random_data, int n_acc, long long* cycles) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const unsigned base = (s32(smem_raw) + 1023u) & ~1023u;
    __shared__ uint64_t bar;
    __shared__ unsigned tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 160 * 1024 / 4; i += 128) reinterpret_cast<float*>(smem_raw + (base - s32(smem_raw)))[i] = random_data ? (float)((int)((i * 2654435761u) >> 20) - 2048) * 0.001f : 1.0f;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = tmem_slot;
    if (tid == 0) {
        const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)mn_major << 15) | ((unsigned)mn_major << 16) |
                               ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
        uint64_t da[4], db[4];
        for (int w = 0; w < 4; ++w) {
            // four different operand windows, so the operands are really re-fetched
            if (mn_major) {
                da[w] = mk_desc(base + (unsigned)w * 1024u + (unsigned)a_shift, (unsigned)a_lbo, 512u, 1u);
                db[w] = mk_desc(base + 65536u + (unsigned)w * 1024u, 16384u, 512u, 1u);
            } else {
                da[w] = mk_desc(base + (unsigned)w * 32u, 0u, 1024u, 2u);
                db[w] = mk_desc(base + 65536u + (unsigned)w * 32u, 0u, 1024u, 2u);
            }
        }
        const long long t0 = clock64();
        for (int it = 0; it < iters; it += 4) {
#pragma unroll
            for (int w = 0; w < 4; ++w)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem + (unsigned)((it + w) % n_acc) * (n_acc > 2 ? 64u : 256u)), "l"(da[w]), "l"(db[w]), "r"(idesc), "r"(it > 0 ? 1u : 0u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
        unsigned done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(s32(&bar)), "r"(0u) : "memory");
        cycles[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    long long* d;
    cudaMalloc(&d, sms * sizeof(long long));
    cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 4096;
    const int shapes[][2] = {{128, 64}, {64, 64}, {128, 128}, {64, 128}, {128, 192}, {128, 256}, {64, 256}, {128, 16}};
    printf("%-10s %-34s %12s %14s %16s\n", "M x N x 8", "operands", "clk / MMA", "MAC / clk / SM", "smem B / clk");
    auto run = [&](int M, int N, int mn, int a_shift, int a_lbo, const char* label, int random_data = 0, int n_acc = 2) {
        k_rate<<<sms, 128, 200 * 1024>>>(M, N, mn, iters, a_shift, a_lbo, random_data, n_acc, d);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); exit(1); }
        long long h[256];
        cudaMemcpy(h, d, sms * sizeof(long long), cudaMemcpyDeviceToHost);
        double mean = 0;
        for (int i = 0; i < sms; ++i) mean += (double)h[i];
        mean /= sms;
        const double clk = mean / iters;
        printf("%3dx%3dx8  %-34s %12.1f %14.0f %16.1f\n", M, N, label, clk, M * N * 8.0 / clk, (M + N) * 32.0 / clk);
    };
    for (int mn = 0; mn < 2; ++mn)
        for (auto& sh : shapes) run(sh[0], sh[1], mn, 0, 16384, mn ? "MN-major" : "K-major");
    // MN-major A whose start address is moved by whole 128-byte rows / whose four M atoms are one row apart
    run(128, 64, 1, 128, 16384, "MN-major, A start + 1 row");
    run(128, 64, 1, 256, 16384, "MN-major, A start + 2 rows");
    run(128, 64, 1, 512, 16384, "MN-major, A start + 4 rows");
    run(128, 64, 1, 0, 128, "MN-major, A atoms 1 row apart");
    run(128, 64, 1, 0, 512, "MN-major, A atoms 4 rows apart");
    run(128, 128, 1, 128, 16384, "MN-major, A start + 1 row");
    run(128, 128, 1, 0, 128, "MN-major, A atoms 1 row apart");
    run(128, 64, 1, 0, 128, "MN-major, random data", 1, 2);
    run(128, 64, 1, 0, 128, "MN-major, random data, 6 accum.", 1, 6);
    run(128, 64, 1, 0, 128, "MN-major, ones, 6 accumulators", 0, 6);
    run(128, 128, 0, 0, 16384, "K-major, random data", 1, 2);
    run(128, 256, 0, 0, 16384, "K-major, random data", 1, 2);
    return 0;
}
