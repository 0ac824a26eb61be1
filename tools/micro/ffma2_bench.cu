// Micro-benchmark: FFMA vs packed FFMA2 (fma.rn.f32x2) issue rate on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 r, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

constexpr int ITER = 4096, CH = 8;
// MODE 0: scalar FFMA, immediate addend; 1: scalar FFMA, register addend; 2: FFMA2 imm; 3: FFMA2 reg;
// 4: FFMA2 imm + 1 LOP3 per FFMA2 (ALU co-issue); 5: scalar FFMA imm + 1 LOP3 per FFMA
template <int MODE>
__global__ void k(float* out, float seed) {
    float t = seed + threadIdx.x * 1e-9f;
    float a[CH], b[CH];
    unsigned m[CH];
    for (int i = 0; i < CH; ++i) { a[i] = t + i; b[i] = t - i; m[i] = threadIdx.x + i; }
    u64 T = pk(t, t * 0.5f), P[CH];
    for (int i = 0; i < CH; ++i) P[i] = pk(a[i], b[i]);
    float r = seed * 3.f;
    u64 R = pk(r, r);
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (MODE == 0) { a[i] = fma1(a[i], t, 1.25f); b[i] = fma1(b[i], t, -0.75f); }
            if (MODE == 1) { a[i] = fma1(a[i], t, r); b[i] = fma1(b[i], t, r); }
            if (MODE == 2) { P[i] = fma2(P[i], T, pk(1.25f, 1.25f)); }
            if (MODE == 3) { P[i] = fma2(P[i], T, R); }
            if (MODE == 4) { P[i] = fma2(P[i], T, pk(1.25f, 1.25f)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(m[i]) : "r"(m[(i + 1) % CH]), "r"(m[(i + 2) % CH])); }
            if (MODE == 5) { a[i] = fma1(a[i], t, 1.25f); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(m[i]) : "r"(m[(i + 1) % CH]), "r"(m[(i + 2) % CH])); }
        }
    }
    float s = 0; unsigned ms = 0;
    for (int i = 0; i < CH; ++i) { float x, y; upk(P[i], x, y); s += a[i] + b[i] + x + y; ms ^= m[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + ms;
}

template <int MODE>
void run(const char* name, double fma_per_thread_iter) {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(out, 1e-3f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) k<MODE><<<148 * 8, 256>>>(out, 1e-3f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    double fmas = 148.0 * 8 * 256 * ITER * CH * fma_per_thread_iter;
    printf("%-40s %8.3f ms  %7.2f TFLOP/s (2*FMA)  err=%s\n", name, ms, 2 * fmas / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}
int main() {
    run<0>("FFMA imm addend", 2);
    run<1>("FFMA reg addend", 2);
    run<2>("FFMA2 imm addend", 2);
    run<3>("FFMA2 reg addend", 2);
    run<4>("FFMA2 imm + LOP3 1:1", 2);
    run<5>("FFMA imm + LOP3 1:1", 1);
    return 0;
}
