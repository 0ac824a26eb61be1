// Micro-benchmark: legacy warp-level mma.sync m16n8k8 TF32 throughput on sm_100a (register operands, no memory).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
constexpr int ITER = 4096, NACC = 8;
__global__ void k(float* out) {
    float c[NACC][4] = {};
    unsigned a[4] = {threadIdx.x, threadIdx.x + 1, threadIdx.x + 2, threadIdx.x + 3}, b[2] = {threadIdx.x * 3, threadIdx.x * 5};
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) mma_tf32(c[i], a, b);
    }
    float s = 0;
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 4 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<148 * 4, 256>>>(out); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) k<<<148 * 4, 256>>>(out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    double flops = 148.0 * 4 * 8 /*warps*/ * ITER * NACC * (16.0 * 8 * 8 * 2);
    printf("mma.sync m16n8k8 tf32: %.3f ms, %.1f TFLOP/s (%s)\n", ms, flops / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
