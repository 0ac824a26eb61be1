// Weight / bias gradient of a per-voxel Dense layer of the amortization network (reference model.py:122-223,
// 1x1x1 convolutions):  dW[o, i] = sum_v g[v, o] * x[v, i],  db[o] = sum_v g[v, o],  N = 10^5..10^6 voxels,
// o, i <= 64.  A reduction GEMM with K = voxels: 2 * 64 * 4 B read per voxel for 8 kFLOP, i.e. HBM-bound at
// 16 FLOP/B -- cuBLAS serves this shape with an sm_80 64x64 kernel at ~270 us for 524 288 voxels where the HBM
// time is ~40 us.  Here a persistent CTA streams 32-voxel tiles of g and x through shared memory (padded rows,
// conflict-free fragment loads), accumulates its 64 x 64 partial with mma.sync m16n8k8 TF32 (fp32 accumulate; the
// tensor pipe is far from the bound, so the warp-level MMA is enough to sit on the HBM roofline), a ones-column
// appended to x yields db from the same MMAs, and a second kernel sums the per-CTA partials in a fixed order.
#include "launch.h"

namespace qb {

namespace {

constexpr int kWgTile = 32;            // voxels per tile (2 buffers x 2 operands x 9 KB of static shared memory)
constexpr int kWgCtasPerSm = 4;
constexpr int kWgLd = 72;              // padded row (floats): bank = (8 t + g) mod 32 for the fragment pattern
constexpr int kWgThreads = 256;

__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
        "{%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

}  // namespace

// g [n, n_out], x [n, n_in] row-major.  partial [gridDim.x, 64, 64]: partial[b][o][i] (column n_in = bias gradient).
__global__ void __launch_bounds__(kWgThreads) k_dense_wgrad(const float* __restrict__ g,
                                                            const float* __restrict__ relu_mask, int n_out,
                                                            const float* __restrict__ x, int n_in, int64_t n,
                                                            float* __restrict__ partial) {
    __shared__ float sg[2][kWgTile * kWgLd];
    __shared__ float sx[2][kWgTile * kWgLd];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gid = lane >> 2, t4 = lane & 3;
    const int m0 = (warp & 1) * 32, n0 = (warp >> 1) * 16;      // this warp's 32 (o) x 16 (i) block of dW
    float acc[2][2][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;

    const int64_t tiles = (n + kWgTile - 1) / kWgTile;
    // Padding is written once: columns >= n_out of g and > n_in of x stay 0, column n_in of x stays 1 (bias
    // gradient); a tile only rewrites the live columns (16-byte cp.async when the row length allows, else scalar).
    for (int e = tid; e < 2 * kWgTile * kWgLd; e += kWgThreads) {
        (&sg[0][0])[e] = 0.f;
        (&sx[0][0])[e] = ((e % kWgLd) == n_in) ? 1.0f : 0.f;
    }
    __syncthreads();
    const bool vec_g = (n_out & 3) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0;
    const bool vec_x = (n_in & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
    auto stage_one = [&](float* dst, const float* __restrict__ src, int ncol, bool vec, int64_t v0, bool ones) {
        const int rows = (int)((n - v0 < kWgTile) ? (n - v0) : kWgTile);
        if (!ones && relu_mask != nullptr) {                          // g * [relu output > 0]: fuses ReLU' (model.py:152)
            if (vec && (reinterpret_cast<uintptr_t>(relu_mask) & 15) == 0) {
                const int per_row = ncol >> 2;
                for (int e = tid; e < kWgTile * per_row; e += kWgThreads) {
                    const int r = e / per_row, c4 = e - r * per_row;
                    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r < rows) {
                        const int64_t at = (v0 + r) * ncol + c4 * 4;
                        const float4 a = __ldg(reinterpret_cast<const float4*>(src + at));
                        const float4 m = __ldg(reinterpret_cast<const float4*>(relu_mask + at));
                        o = make_float4(m.x > 0.f ? a.x : 0.f, m.y > 0.f ? a.y : 0.f, m.z > 0.f ? a.z : 0.f,
                                        m.w > 0.f ? a.w : 0.f);
                    }
                    *reinterpret_cast<float4*>(dst + r * kWgLd + c4 * 4) = o;
                }
            } else {
                for (int e = tid; e < kWgTile * ncol; e += kWgThreads) {
                    const int r = e / ncol, c = e - r * ncol;
                    const int64_t at = (v0 + r) * ncol + c;
                    dst[r * kWgLd + c] = (r < rows && __ldg(relu_mask + at) > 0.f) ? __ldg(src + at) : 0.f;
                }
            }
        } else if (vec) {
            const int per_row = ncol >> 2;
            for (int e = tid; e < kWgTile * per_row; e += kWgThreads) {
                const int r = e / per_row, c4 = e - r * per_row;
                float* d = dst + r * kWgLd + c4 * 4;
                if (r < rows) {
                    const unsigned sa = (unsigned)__cvta_generic_to_shared(d);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src + (v0 + r) * ncol + c4 * 4)
                                 : "memory");
                } else {
                    *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        } else {
            for (int e = tid; e < kWgTile * ncol; e += kWgThreads) {
                const int r = e / ncol, c = e - r * ncol;
                dst[r * kWgLd + c] = (r < rows) ? __ldg(src + (v0 + r) * ncol + c) : 0.f;
            }
        }
        if (ones && rows < kWgTile)                                   // tail tile: rows beyond n contribute nothing
            for (int r = rows + tid; r < kWgTile; r += kWgThreads) dst[r * kWgLd + ncol] = 0.f;
    };
    auto stage = [&](int b, int64_t tile_idx) {
        stage_one(sg[b], g, n_out, vec_g, tile_idx * kWgTile, false);
        stage_one(sx[b], x, n_in, vec_x, tile_idx * kWgTile, true);
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int buf = 0;
    int64_t tile = blockIdx.x;
    if (tile < tiles) stage(0, tile);
    for (; tile < tiles; tile += gridDim.x) {
        const int64_t next = tile + gridDim.x;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                                             // tile `buf` landed; buffer buf^1 is free again
        if (next < tiles) stage(buf ^ 1, next);                      // in flight while the MMAs below run
        const float* G = sg[buf];
        const float* X = sx[buf];
#pragma unroll
        for (int k0 = 0; k0 < kWgTile; k0 += 8) {
            unsigned a[2][4], b[2][2];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {                     // A[m = o][k = v] = g[v][o]
                const int o = m0 + mt * 16 + gid;
                a[mt][0] = __float_as_uint(G[(k0 + t4) * kWgLd + o]);
                a[mt][1] = __float_as_uint(G[(k0 + t4) * kWgLd + o + 8]);
                a[mt][2] = __float_as_uint(G[(k0 + t4 + 4) * kWgLd + o]);
                a[mt][3] = __float_as_uint(G[(k0 + t4 + 4) * kWgLd + o + 8]);
            }
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {                     // B[k = v][n = i] = x[v][i]
                const int i = n0 + nt * 8 + gid;
                b[nt][0] = __float_as_uint(X[(k0 + t4) * kWgLd + i]);
                b[nt][1] = __float_as_uint(X[(k0 + t4 + 4) * kWgLd + i]);
            }
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) mma_tf32(acc[mt][nt], a[mt], b[nt]);
        }
        buf ^= 1;
    }
    float* out = partial + (int64_t)blockIdx.x * 64 * 64;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            const int o = m0 + mt * 16 + gid, i = n0 + nt * 8 + 2 * t4;
            out[o * 64 + i] = acc[mt][nt][0];
            out[o * 64 + i + 1] = acc[mt][nt][1];
            out[(o + 8) * 64 + i] = acc[mt][nt][2];
            out[(o + 8) * 64 + i + 1] = acc[mt][nt][3];
        }
}

// dW [n_out, n_in] (+)= sum_b partial[b][o][i];  db [n_out] (+)= sum_b partial[b][o][n_in]; fixed summation order.
__global__ void k_dense_wgrad_reduce(const float* __restrict__ partial, int n_parts, int n_out, int n_in,
                                     float* __restrict__ dw, float* __restrict__ db, int accumulate) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 64 * 64) return;
    const int o = e >> 6, i = e & 63;
    if (o >= n_out || i > n_in) return;
    float s = 0.f;
    for (int b = 0; b < n_parts; ++b) s += partial[(int64_t)b * 4096 + e];
    if (i < n_in) {
        float* d = dw + o * n_in + i;
        *d = accumulate ? *d + s : s;
    } else if (db != nullptr) {
        db[o] = accumulate ? db[o] + s : s;
    }
}

}  // namespace qb

using namespace qb;

extern "C" int64_t qbold_dense_wgrad_workspace_floats(void) { return (int64_t)sm_count() * kWgCtasPerSm * 64 * 64; }

extern "C" int qbold_dense_wgrad(const float* g, const float* relu_mask, int32_t n_out, const float* x, int32_t n_in,
                                 int64_t n, float* dw, float* db, int32_t accumulate, float* workspace, void* stream) {
    if (n_out < 1 || n_out > 64 || n_in < 1 || n_in > 63 || n < 0)
        return fail(QBOLD_EUNSUPPORTED, "qbold_dense_wgrad: supports n_out <= 64, n_in <= 63 (got %d, %d)", n_out, n_in);
    if (!g || !x || !dw || !workspace) return fail(QBOLD_EINVAL, "qbold_dense_wgrad: null pointer");
    const int64_t tiles = (n + kWgTile - 1) / kWgTile;
    int64_t grid = (int64_t)sm_count() * kWgCtasPerSm;
    if (tiles < grid) grid = tiles;
    if (grid < 1) grid = 1;
    k_dense_wgrad<<<(unsigned)grid, kWgThreads, 0, (cudaStream_t)stream>>>(g, relu_mask, n_out, x, n_in, n, workspace);
    int rc = after_launch("k_dense_wgrad");
    if (rc) return rc;
    k_dense_wgrad_reduce<<<16, 256, 0, (cudaStream_t)stream>>>(workspace, (int)grid, n_out, n_in, dw, db, accumulate);
    return after_launch("k_dense_wgrad_reduce");
}

// ---- skinny Dense layers (the encoder's heads: 60 -> 5 posterior parameters, 60 -> 11 sigmas) --------------------------
// The skinny Dense kernels (n_out <= 16) live in encoder_small_kernels.cuh; their launchers follow.
#include "encoder_small_kernels.cuh"

static int small_dense_args_ok(const void* a, const void* w, const void* out, int n_in, int n_out, int64_t n) {
    return a && w && out && n >= 0 && n_out >= 1 && n_out <= kSmallOut && n_in >= 4 && n_in <= kSmallIn && (n_in & 3) == 0 &&
           (reinterpret_cast<uintptr_t>(a) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0 &&
           (reinterpret_cast<uintptr_t>(out) & 15) == 0;
}

extern "C" int qbold_dense_small_forward(const float* x, const float* w, const float* bias, int32_t n_in, int32_t n_out,
                                         int64_t n, float* y, void* stream) {
    if (!small_dense_args_ok(x, w, y, n_in, n_out, n) || !bias)
        return fail(QBOLD_EUNSUPPORTED, "qbold_dense_small_forward: needs n_out <= 16, n_in a multiple of 4 up to 64, "
                                        "16-byte aligned x / w / y");
    if (n == 0) return QBOLD_OK;
    int64_t grid = (n + 127) / 128;
    const int64_t cap = (int64_t)sm_count() * 6;
    if (grid > cap) grid = cap;
    k_dense_small_fwd_coop<<<(unsigned)grid, 128, 0, (cudaStream_t)stream>>>(x, w, bias, n_in, n_out, n, y);
    return after_launch("k_dense_small_fwd_coop");
}

extern "C" int qbold_dense_small_dgrad(const float* g, const float* w, int32_t n_in, int32_t n_out, int64_t n, float* dx,
                                       void* stream) {
    if (!small_dense_args_ok(dx, w, dx, n_in, n_out, n) || !g)
        return fail(QBOLD_EUNSUPPORTED, "qbold_dense_small_dgrad: needs n_out <= 16, n_in a multiple of 4 up to 64, "
                                        "16-byte aligned w / dx");
    if (n == 0) return QBOLD_OK;
    int64_t grid = (n + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (grid > cap) grid = cap;
    k_dense_small_dgrad_coop<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(g, w, nullptr, n_in, n_out, n, dx);
    return after_launch("k_dense_small_dgrad_coop");
}

extern "C" int qbold_dense_small_dgrad_masked(const float* g, const float* w, const float* relu_mask, int32_t n_in,
                                              int32_t n_out, int64_t n, float* dx, void* stream) {
    if (!g || !w || !dx || n < 0 || n_out < 1 || n_out > kSmallOut || n_in < 1 || n_in > kSmallIn)
        return fail(QBOLD_EUNSUPPORTED, "qbold_dense_small_dgrad_masked: needs n_out <= 16, n_in <= 64");
    if (n == 0) return QBOLD_OK;
    int64_t grid = (n + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (grid > cap) grid = cap;
    k_dense_small_dgrad_coop<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(g, w, relu_mask, n_in, n_out, n, dx);
    return after_launch("k_dense_small_dgrad_coop");
}
