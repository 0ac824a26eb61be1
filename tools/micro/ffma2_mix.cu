// Micro-benchmark: does packed FFMA2 free issue slots for other pipes (ALU / LSU / XU) on sm_100a?
// Every variant does the same number of FMAs per thread; "x" instructions are added per 8 FMAs.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 r, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

constexpr int ITER = 4096;
// PACK: 0 = 8 scalar FFMA per iteration, 1 = 4 FFMA2.  KIND: 0 none, 1 LOP3, 2 LDS, 3 MUFU.RSQ, 4 IADD3, 5 FMNMX  NX: extra instr per iteration
template <int PACK, int KIND, int NX>
__global__ void k(float* out, float seed) {
    __shared__ float sm[256 * 4];
    for (int i = 0; i < 4; ++i) sm[threadIdx.x * 4 + i] = seed * i;
    __syncthreads();
    float t = seed + threadIdx.x * 1e-9f;
    float a[8];
    unsigned m[8];
    float f[8];
    for (int i = 0; i < 8; ++i) { a[i] = t + i; m[i] = threadIdx.x * 7 + i; f[i] = t * i + 1.0f; }
    u64 T = pk(t, t * 0.5f), P[4];
    for (int i = 0; i < 4; ++i) P[i] = pk(a[2 * i], a[2 * i + 1]);
    unsigned k1 = threadIdx.x | 0x55, k2 = threadIdx.x * 3;
    unsigned sa = (unsigned)__cvta_generic_to_shared(&sm[threadIdx.x]);
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (PACK == 0) a[i] = fma1(a[i], t, 1.25f);
            if (PACK == 1 && (i & 1)) P[i >> 1] = fma2(P[i >> 1], T, pk(1.25f, 1.25f));
            if (i < NX) {
                if (KIND == 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(m[i]) : "r"(k1), "r"(k2));
                if (KIND == 2) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(f[i]) : "r"(sa + i * 4 * 32));
                if (KIND == 3) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
                if (KIND == 4) asm volatile("add.s32 %0, %0, %1;" : "+r"(m[i]) : "r"(k1));
                if (KIND == 5) asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(t));
            }
        }
    }
    float s = 0; unsigned ms = 0;
    for (int i = 0; i < 8; ++i) { s += a[i] + f[i]; ms ^= m[i]; }
    for (int i = 0; i < 4; ++i) { float x, y; upk(P[i], x, y); s += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + ms;
}

template <int PACK, int KIND, int NX>
void run(const char* name) {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<PACK, KIND, NX><<<148 * 8, 256>>>(out, 1e-3f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) k<PACK, KIND, NX><<<148 * 8, 256>>>(out, 1e-3f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    // cycles per SMSP per iteration-of-8-FMAs: 16 warps per SMSP
    double cyc = ms * 1e-3 * 1.965e9 / (16.0 * ITER);
    printf("%-28s pack=%d extra=%d  %7.3f ms  %5.2f cyc/(8 FMA + extras) per warp-slot  %s\n", name, PACK, NX, ms, cyc, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}
#define BOTH(K, N, name) run<0, K, N>(name); run<1, K, N>(name);
int main() {
    BOTH(0, 0, "fma only");
    BOTH(1, 2, "LOP3"); BOTH(1, 4, "LOP3"); BOTH(1, 8, "LOP3");
    BOTH(4, 2, "IADD"); BOTH(4, 4, "IADD"); BOTH(4, 8, "IADD");
    BOTH(5, 4, "FMNMX");
    BOTH(2, 2, "LDS"); BOTH(2, 4, "LDS");
    BOTH(3, 1, "MUFU.RSQ"); BOTH(3, 2, "MUFU.RSQ");
    return 0;
}
