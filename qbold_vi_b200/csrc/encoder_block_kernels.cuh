// Streaming kernels of the fused encoder block (reference create_block, model.py:142-174) used by the training
// path of qbold_vi_b200/encoder.py (_BlockFn).  The block's convolutions and 60x60 GEMMs run in the libraries; every
// elementwise step between them is fused here so that an activation tensor ([voxels, 60] floats, 126 MB for
// 2 x 64^3 volumes) is read and written as few times as the data flow allows:
//
//   k_block_mix_fwd    out = skip (1 - g) + (r0 + b_r) g,  g = sigmoid(z + offset)     (model.py:160-172)
//                      r0 is the second 3x3x1 convolution WITHOUT its bias (cuDNN adds a bias in a separate pass);
//                      optionally also writes relu(out), the next block's convolution input (model.py:150).
//   k_block_mix_bwd    d_r, d_z and d_skip * [skip > 0] (the skip branch ends in a ReLU: its derivative is applied
//                      here instead of in a pass of its own).
//   k_relu_bwd_colsum  g * [y > 0] and, in the same pass, its column sums (= the bias gradient of the layer that
//                      produced y); deterministic two-stage reduction.
//   k_colsum           column sums alone (bias gradient of a convolution without ReLU).
// All HBM-bound: 16-byte accesses, grid = a multiple of the SM count.
//
// The kernels live in this header, apart from their launchers in encoder_block.cu, so that tests/host_emu can compile
// the same kernel source for the host and run it in its SIMT emulator (CPU suite).
#pragma once
#include "launch.h"

namespace qb {

namespace {

constexpr int kColTile = 256;                  // threads per CTA of the column-sum kernels
constexpr int kMaxC4 = 16;                     // up to 64 channels (16 float4 per row)

__device__ __forceinline__ float gate(float z, float offset) { return 1.0f / (1.0f + expf(-(z + offset))); }

__device__ __forceinline__ float4 ld4(const float4* p) { return __ldg(p); }

}  // namespace

__global__ void __launch_bounds__(kThreads) k_block_mix_fwd(const float4* __restrict__ skip, const float4* __restrict__ r0,
                                                            const float4* __restrict__ r_bias,
                                                            const float4* __restrict__ z, float offset, int64_t total4,
                                                            int c4, float4* __restrict__ out,
                                                            float4* __restrict__ out_relu) {
    for (int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x; e < total4; e += (int64_t)gridDim.x * kThreads) {
        const float4 zz = ld4(z + e), s = ld4(skip + e);
        float4 rr = ld4(r0 + e);
        if (r_bias != nullptr) {
            const float4 b = ld4(r_bias + (int)(e % c4));
            rr = make_float4(rr.x + b.x, rr.y + b.y, rr.z + b.z, rr.w + b.w);
        }
        const float g0 = gate(zz.x, offset), g1 = gate(zz.y, offset), g2 = gate(zz.z, offset), g3 = gate(zz.w, offset);
        const float4 o = make_float4(s.x * (1.0f - g0) + rr.x * g0, s.y * (1.0f - g1) + rr.y * g1,
                                     s.z * (1.0f - g2) + rr.z * g2, s.w * (1.0f - g3) + rr.w * g3);
        out[e] = o;
        if (out_relu != nullptr)
            out_relu[e] = make_float4(fmaxf(o.x, 0.f), fmaxf(o.y, 0.f), fmaxf(o.z, 0.f), fmaxf(o.w, 0.f));
    }
}

__global__ void __launch_bounds__(kThreads) k_block_mix_bwd(const float4* __restrict__ go, const float4* __restrict__ skip,
                                                            const float4* __restrict__ r0,
                                                            const float4* __restrict__ r_bias,
                                                            const float4* __restrict__ z, float offset, int64_t total4,
                                                            int c4, int skip_is_relu, const float4* __restrict__ skip_addend, float4* __restrict__ d_skip,
                                                            float4* __restrict__ d_r, float4* __restrict__ d_z) {
    for (int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x; e < total4; e += (int64_t)gridDim.x * kThreads) {
        const float4 zz = ld4(z + e), s = ld4(skip + e), o = ld4(go + e);
        float4 rr = ld4(r0 + e);
        if (r_bias != nullptr) {
            const float4 b = ld4(r_bias + (int)(e % c4));
            rr = make_float4(rr.x + b.x, rr.y + b.y, rr.z + b.z, rr.w + b.w);
        }
        const float g0 = gate(zz.x, offset), g1 = gate(zz.y, offset), g2 = gate(zz.z, offset), g3 = gate(zz.w, offset);
        float4 ds = make_float4(o.x * (1.0f - g0), o.y * (1.0f - g1), o.z * (1.0f - g2), o.w * (1.0f - g3));
        if (skip_addend != nullptr) {                                // a second gradient of the same activation (stream 1)
            const float4 a = ld4(skip_addend + e);
            ds = make_float4(ds.x + a.x, ds.y + a.y, ds.z + a.z, ds.w + a.w);
        }
        if (skip_is_relu)                                            // skip = relu(.): threshold_backward(ds, skip, 0)
            ds = make_float4(s.x > 0.f ? ds.x : 0.f, s.y > 0.f ? ds.y : 0.f, s.z > 0.f ? ds.z : 0.f,
                             s.w > 0.f ? ds.w : 0.f);
        d_skip[e] = ds;
        d_r[e] = make_float4(o.x * g0, o.y * g1, o.z * g2, o.w * g3);
        d_z[e] = make_float4(o.x * (rr.x - s.x) * (g0 * (1.0f - g0)), o.y * (rr.y - s.y) * (g1 * (1.0f - g1)),
                             o.z * (rr.z - s.z) * (g2 * (1.0f - g2)), o.w * (rr.w - s.w) * (g3 * (1.0f - g3)));
    }
}

// Rows [row0, row1) of this CTA: thread (lane-in-row-group) owns one float4 column chunk of every (kColTile / c4)-th
// row.  MASK: out = g * [y > 0] is written as well.  partial [gridDim.x, 4 c4] receives the CTA's column sums.
template <bool MASK>
__global__ void __launch_bounds__(kColTile) k_relu_bwd_colsum(const float4* __restrict__ g, const float4* __restrict__ y,
                                                              const float4* __restrict__ addend, int64_t n, int c4,
                                                              float4* __restrict__ out, float* __restrict__ partial) {
    __shared__ float4 red[kColTile];
    const int rows_per_pass = kColTile / c4;                     // threads beyond rows_per_pass * c4 idle
    const int rr = threadIdx.x / c4, cc = threadIdx.x % c4;
    const bool active = rr < rows_per_pass;
    const int64_t rows_per_cta = (n + gridDim.x - 1) / gridDim.x;
    const int64_t row0 = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t row1 = row0 + rows_per_cta < n ? row0 + rows_per_cta : n;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) {
        for (int64_t row = row0 + rr; row < row1; row += rows_per_pass) {
            const int64_t e = row * c4 + cc;
            float4 v = ld4(g + e);
            if (MASK) {
                const float4 m = ld4(y + e);
                v = make_float4(m.x > 0.f ? v.x : 0.f, m.y > 0.f ? v.y : 0.f, m.z > 0.f ? v.z : 0.f,
                                m.w > 0.f ? v.w : 0.f);
                if (addend != nullptr) {
                    const float4 a = ld4(addend + e);
                    v = make_float4(v.x + a.x, v.y + a.y, v.z + a.z, v.w + a.w);
                }
                if (out != nullptr) out[e] = v;
            }
            acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
        }
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    if (partial != nullptr && threadIdx.x < c4) {                // fixed order: deterministic
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < rows_per_pass; ++r) {
            const float4 v = red[r * c4 + threadIdx.x];
            s.x += v.x, s.y += v.y, s.z += v.z, s.w += v.w;
        }
        reinterpret_cast<float4*>(partial)[(int64_t)blockIdx.x * c4 + threadIdx.x] = s;
    }
}

// Second stage: column j of the [n_parts, c] partial sums, summed in a FIXED order (16 strided slices per column, then
// the 16 slice totals in order) so the result does not depend on scheduling.  One CTA of 16 x 64 threads.
__global__ void __launch_bounds__(1024) k_colsum_finish(const float* __restrict__ partial, int n_parts, int c,
                                                        float* __restrict__ out, int accumulate) {
    __shared__ float red[16][64];
    const int j = threadIdx.x & 63, slice = threadIdx.x >> 6;
    float s = 0.f;
    if (j < c)
        for (int p = slice; p < n_parts; p += 16) s += partial[(int64_t)p * c + j];
    red[slice][j] = s;
    __syncthreads();
    if (slice == 0 && j < c) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) t += red[k][j];
        out[j] = accumulate ? out[j] + t : t;
    }
}

// normalise_data (model.py:97-113) fused with the layout change of the training path: raw images [B, X, Y, Z, T] ->
// log(clip(d, 1e-2, 1e8) / reference) as z-outer rows [B, Z, X, Y, Tp], Tp = T rounded up to a multiple of 4 (zero
// padded, so the rows are 16-byte aligned operands of the TMA-fed first Dense layer).  A CTA transposes a 32 (y) x 32 (z)
// tile through shared memory: reads run along z (the input's fastest spatial axis), writes along y (the output's).
__global__ void __launch_bounds__(256) k_normalise_zouter(const float* __restrict__ data, int X, int Y, int Z, int T,
                                                         int Tp, int se, int multi, float* __restrict__ out) {
#ifdef QB_HOST_EMU
    float* tile = qb_emu::dynamic_smem();                                   // tests/host_emu: the launch's dynamic window
#else
    extern __shared__ float tile[];                                         // [32 y][32 T + 1]
#endif
    const int zt = (Z + 31) / 32, yt = (Y + 31) / 32;
    int blk = blockIdx.x;
    const int z0 = (blk % zt) * 32;
    blk /= zt;
    const int y0 = (blk % yt) * 32;
    blk /= yt;
    const int x = blk % X, b = blk / X;
    const int nz = min(32, Z - z0), ny = min(32, Y - y0);
    const int run = nz * T, pitch = 32 * T + 1;
    for (int e = threadIdx.x; e < ny * run; e += blockDim.x) {
        const int y = e / run, r = e - y * run;
        tile[y * pitch + r] = __ldg(data + ((((int64_t)b * X + x) * Y + (y0 + y)) * Z + z0) * T + r);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane >= ny) return;
    for (int z = warp; z < nz; z += 8) {
        const float* src = tile + lane * pitch + z * T;
        float ref;
        if (multi) {
            ref = (fminf(fmaxf(src[se - 1], 1e-2f), 1e8f) + fminf(fmaxf(src[se], 1e-2f), 1e8f) +
                   fminf(fmaxf(src[se + 1], 1e-2f), 1e8f)) / 3.0f;
        } else {
            ref = fminf(fmaxf(src[se], 1e-2f), 1e8f);
        }
        float4* dst = reinterpret_cast<float4*>(out + ((((int64_t)b * Z + (z0 + z)) * X + x) * Y + (y0 + lane)) * Tp);
        for (int t = 0; t < Tp; t += 4) {                                   // 16-byte stores (Tp is a multiple of 4)
            float4 v;
            v.x = logf(fminf(fmaxf(src[t], 1e-2f), 1e8f) / ref);
            v.y = t + 1 < T ? logf(fminf(fmaxf(src[t + 1], 1e-2f), 1e8f) / ref) : 0.f;
            v.z = t + 2 < T ? logf(fminf(fmaxf(src[t + 2], 1e-2f), 1e8f) / ref) : 0.f;
            v.w = t + 3 < T ? logf(fminf(fmaxf(src[t + 3], 1e-2f), 1e8f) / ref) : 0.f;
            dst[t >> 2] = v;
        }
    }
}

}  // namespace qb
