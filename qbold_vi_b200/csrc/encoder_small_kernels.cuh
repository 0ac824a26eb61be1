// Skinny Dense layers of the encoder's heads (60 -> 5, 60 -> 11; reference create_encoder, model.py:176-223) as streaming
// kernels.  Kept in a header, apart from their launchers in encoder_wgrad.cu, so that tests/host_emu can compile the same
// kernel source for the host and run it in its SIMT emulator (CPU suite).
#pragma once
#include "launch.h"

// y[v, o] = b[o] + sum_i x[v, i] W[o, i]  and  dx[v, i] = sum_o g[v, o] W[o, i]  for n_out <= 16.  With so few outputs the
// GEMM is a streaming pass over x (or dx): one thread per voxel, W in shared memory read as broadcast float4, ~25 us of
// HBM time per 524 288 voxels where the library's tensor-op kernels take 100-130 us on this shape.
namespace qb {

constexpr int kSmallOut = 16, kSmallIn = 64;

// Forward with coalesced loads: a warp copies 32 consecutive rows of x (one contiguous 32 * n_in float run) into its
// shared-memory tile with a row pitch of n_in + 1 floats, then lane l multiplies row l (bank-conflict free: the pitch is
// odd) with the weights broadcast from shared memory.  58 -> ~25 us for the 60 -> 5 head on 524 288 voxels.
__global__ void __launch_bounds__(128) k_dense_small_fwd_coop(const float* __restrict__ x, const float* __restrict__ w,
                                                              const float* __restrict__ b, int n_in, int n_out, int64_t n,
                                                              float* __restrict__ y) {
    __shared__ float sw[kSmallOut][kSmallIn];
    __shared__ float sb[kSmallOut];
    __shared__ float tile[4][32 * (kSmallIn + 1)];
    for (int e = threadIdx.x; e < kSmallOut * kSmallIn; e += blockDim.x) {
        const int o = e / kSmallIn, i = e % kSmallIn;
        sw[o][i] = (o < n_out && i < n_in) ? __ldg(w + o * n_in + i) : 0.f;
    }
    if (threadIdx.x < kSmallOut) sb[threadIdx.x] = threadIdx.x < n_out ? b[threadIdx.x] : 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, pitch = n_in + 1;
    float* t = tile[wid];
    const int64_t warp0 = (int64_t)blockIdx.x * 4 + wid, nwarps = (int64_t)gridDim.x * 4;
    for (int64_t base = warp0 * 32; base < n; base += nwarps * 32) {
        const int cnt = (int)((n - base) < 32 ? (n - base) : 32);
        const int run = cnt * n_in;
        const float* src = x + base * n_in;
        for (int e = lane; e < run; e += 32) {
            const int r = e / n_in, c = e - r * n_in;
            t[r * pitch + c] = __ldg(src + e);
        }
        __syncwarp();
        if (lane < cnt) {
            float acc[kSmallOut];
#pragma unroll
            for (int o = 0; o < kSmallOut; ++o) acc[o] = sb[o];
            const float* row = t + lane * pitch;
            for (int i = 0; i < n_in; ++i) {
                const float a = row[i];
#pragma unroll
                for (int o = 0; o < kSmallOut; ++o)
                    if (o < n_out) acc[o] = fmaf(a, sw[o][i], acc[o]);
            }
#pragma unroll
            for (int o = 0; o < kSmallOut; ++o)
                if (o < n_out) y[(base + lane) * n_out + o] = acc[o];
        }
        __syncwarp();
    }
}

// Input gradient with warp-cooperative, coalesced stores and an optional ReLU' mask on the RESULT:
//     dx[v, i] = [relu_mask[v, i] > 0] * sum_o g[v, o] W[o, i]
// A warp takes 32 voxels: lane l loads the gradient row of voxel l (n_out <= 16 floats), then for each voxel the row
// is broadcast by shuffles and lane l produces inputs i = l and l + 32, so a voxel's 240-byte row is written (and
// its mask row read) by consecutive lanes.  With the mask this also replaces the separate ReLU' pass over the
// activation that the head reads (the last block's stream-1 output).
__global__ void __launch_bounds__(256) k_dense_small_dgrad_coop(const float* __restrict__ g, const float* __restrict__ w,
                                                                const float* __restrict__ relu_mask, int n_in, int n_out,
                                                                int64_t n, float* __restrict__ dx) {
    __shared__ float sw[kSmallOut][kSmallIn];
    for (int e = threadIdx.x; e < kSmallOut * kSmallIn; e += blockDim.x) {
        const int o = e / kSmallIn, i = e % kSmallIn;
        sw[o][i] = (o < n_out && i < n_in) ? __ldg(w + o * n_in + i) : 0.f;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int i0 = lane, i1 = lane + 32;
    float w0[kSmallOut], w1[kSmallOut];
#pragma unroll
    for (int o = 0; o < kSmallOut; ++o) {
        w0[o] = sw[o][i0];
        w1[o] = sw[o][i1];
    }
    for (int64_t base = warp0 * 32; base < n; base += nwarps * 32) {
        const int64_t mine = base + lane;
        float gv[kSmallOut];
#pragma unroll
        for (int o = 0; o < kSmallOut; ++o) gv[o] = (o < n_out && mine < n) ? __ldg(g + mine * n_out + o) : 0.f;
        const int cnt = (int)((n - base) < 32 ? (n - base) : 32);
        for (int t = 0; t < cnt; ++t) {
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int o = 0; o < kSmallOut; ++o) {
                if (o < n_out) {
                    const float gg = __shfl_sync(0xffffffffu, gv[o], t);
                    a0 = fmaf(gg, w0[o], a0);
                    a1 = fmaf(gg, w1[o], a1);
                }
            }
            const int64_t row = (base + t) * n_in;
            if (relu_mask != nullptr) {
                if (i0 < n_in && !(__ldg(relu_mask + row + i0) > 0.f)) a0 = 0.f;
                if (i1 < n_in && !(__ldg(relu_mask + row + i1) > 0.f)) a1 = 0.f;
            }
            if (i0 < n_in) dx[row + i0] = a0;
            if (i1 < n_in) dx[row + i1] = a1;
        }
    }
}

}  // namespace qb
