// Weight gradient of the encoder's 3x3x1 convolutions (reference create_block, model.py:152,156) on the 5th-generation
// tensor cores:
//     dW[o, i, kx, ky] = sum_v g[v, o] * x[v + (kx-1, ky-1), i]          (zero outside the image, padding 'same')
// for activations kept z-outer, [B*Z, X, Y, C] flattened to rows v (qbold_vi_b200/encoder.py), C <= 64.  It is a
// GEMM whose contraction runs over the VOXELS: 34.8 GFLOP on 252 MB for 2 x 64^3 x 60 -- cuDNN's kernel takes 209 us
// (166 TFLOP/s, 60 % of what warp-level mma.sync TF32 can do on this part: tools/micro/mma_tf32_bench.cu), the HBM
// time is 40 us, the tcgen05 time ~35 us.
//
// Mapping.  A persistent CTA (one per SM) walks 32-voxel k-tiles.  For every tile it TRANSPOSES, while staging,
//   * the nine shifted 32-voxel windows of x into five A tiles [128 rows x 32 k]: rows 0-63 = input channels of tap 2t,
//     rows 64-127 = input channels of tap 2t+1 (tap = 3 (dx+1) + (dy+1); the tenth half stays zero), and
//   * the g tile into one B tile [64 rows (output channels) x 32 k],
// both K-major in the canonical SWIZZLE_128B layout (k = voxel within the tile: exactly one 128-byte swizzle row), so
// the operands use the same shared-memory / instruction descriptors as csrc/encoder_mlp.cu.  A lane owns one voxel
// and writes its four channel values to four different rows at the same k: 32 lanes hit 32 different banks.
// Thread 0 then issues 5 x 4 tcgen05.mma kind::tf32 (M = 128, N = 64, K = 8) that accumulate
//     D_t[(tap half, i), o] += sum_k A_t[(tap half, i), k] * B[o, k]
// into five 64-column TMEM accumulators that live for the whole kernel; tcgen05.commit -> mbarrier releases the operand
// buffer for the next tile's transpose, while cp.async already fetches the tile after it.  At the end every warp reads its
// TMEM lane quarter and writes the CTA's partial [5][128][64]; k_conv_wgrad_reduce sums the partials in a fixed order.
#include "launch.h"

namespace qb {

namespace {

constexpr int kCwThreads = 512;
constexpr int kCwTile = 32;                   // voxels per k-tile = floats per 128-byte swizzle row
constexpr int kCwGroups = 5;                  // accumulators: taps (0,1) (2,3) (4,5) (6,7) (8,-)
constexpr int kCwATile = 128 * 128;           // bytes: 128 rows x 128 B
constexpr int kCwBTile = 64 * 128;            // bytes
constexpr int kCwBuf = kCwGroups * kCwATile + kCwBTile;      // 90 112 B: the operand buffer
constexpr int kCwHalo = kCwTile + 2;          // raw rows per x-offset: the tile and one row either side
constexpr int kCwRawMax = (3 * kCwHalo * 68 + kCwTile * 68) * 4;   // bytes of one raw buffer at the widest row stride
constexpr int kCwTmemCols = 512;
constexpr int kCwPartial = kCwGroups * 128 * 64;             // floats per CTA

__device__ __forceinline__ unsigned cw_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cw_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void cw_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void cw_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void cw_mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// Bounded wait (a descriptor mistake must not wedge the GPU); false on timeout.
__device__ __forceinline__ bool cw_mbar_wait(unsigned bar, unsigned parity) {
    for (unsigned spin = 0; spin < (1u << 22); ++spin) {
        unsigned done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return true;
    }
    return false;
}
__device__ __forceinline__ void cw_commit(unsigned bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void cw_mma(unsigned tmem_d, uint64_t desc_a, uint64_t desc_b, unsigned idesc, unsigned acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc)
        : "memory");
}
// K-major SWIZZLE_128B descriptor: start address (16-byte units), 1024 B between 8-row atoms, version 1, layout 2
__device__ __forceinline__ uint64_t cw_desc(unsigned addr) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(1024u >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// D = F32, A = B = TF32, both K-major, M = 128, N = 64
__device__ __forceinline__ unsigned cw_idesc() {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(64 >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
}
__device__ __forceinline__ void cw_tmem_ld16(unsigned taddr, float* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
        "%15}, [%16];"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
          "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// byte offset of element (row, k) in a one-K-block K-major SWIZZLE_128B tile
__device__ __forceinline__ unsigned cw_off(int row, int k) {
    return (unsigned)((row >> 3) * 1024 + (row & 7) * 128 + ((((k >> 2) ^ (row & 7))) << 4) + (k & 3) * 4);
}
__device__ __forceinline__ void cw_sts(unsigned addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

}  // namespace

// g [n, cg], x [n, cx] (z-outer rows v = (image, x, y)); partial [gridDim.x][5][128][64].
//
// Staging in two steps.  (1) cp.async (16-byte, coalesced, L2 -> shared) brings the tile's raw rows in as they lie in
// memory: for each x-offset dx the 34 consecutive rows [v0 + dx Y - 1, v0 + dx Y + 33) of x (a tap's window is a
// constant offset in the flat row index; rows outside the image are masked per lane, never by address), and the 32
// rows of g -- double buffered, so the fetch of tile n+1 overlaps everything else.  (2) shared -> shared transpose
// into the operand tiles: a lane owns one voxel, reads a 16-byte channel chunk of its (shifted) row -- the row stride
// is chosen so that a quarter-warp's 8 LDS.128 fall into disjoint banks -- and writes the four values to four
// operand rows at k = lane, 32 different banks.  (A direct global -> operand transpose has every lane of a load in a
// different 128-byte line and was measured 4x slower than cuDNN.)
__global__ void __launch_bounds__(kCwThreads, 1) k_conv_wgrad_tc(const float* __restrict__ g, int cg,
                                                                  const float* __restrict__ x, int cx, int64_t n,
                                                                  int X, int Y, float* __restrict__ partial,
                                                                  int* __restrict__ status) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const unsigned raw = cw_smem_u32(smem_raw);
    const unsigned base = (raw + 1023u) & ~1023u;                       // operand buffer: 5 A tiles + 1 B tile
    unsigned char* sm = smem_raw + (base - raw);
    const int cx4 = cx >> 2, cg4 = cg >> 2;
    const int ldx = (cx4 & 1) ? cx : cx + 4, ldg_ = (cg4 & 1) ? cg : cg + 4;   // row strides (floats): odd number of 16-byte chunks
    const unsigned raw_bytes = (unsigned)(3 * kCwHalo * ldx + kCwTile * ldg_) * 4u;
    const unsigned raw0 = base + kCwBuf;                                 // raw[b] = raw0 + b * raw_bytes
    uint64_t* sBar = reinterpret_cast<uint64_t*>(sm + kCwBuf + 2 * kCwRawMax);
    unsigned* sTmem = reinterpret_cast<unsigned*>(sBar + 1);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(cw_smem_u32(sTmem)),
                     "r"((unsigned)kCwTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        cw_mbar_init(cw_smem_u32(sBar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {   // padding rows of the operand tiles (channels >= cx / cg, the tenth tap half) are never written: zero them once
        float4* z = reinterpret_cast<float4*>(sm);
        for (int i = tid; i < kCwBuf / 16; i += kCwThreads) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    cw_fence_async();
    cw_before_sync();
    __syncthreads();
    cw_after_sync();
    const unsigned tmem = *sTmem;
    const unsigned idesc = cw_idesc();
    const int n_tasks = 9 * cx4 + cg4;
    const int64_t tiles = (n + kCwTile - 1) / kCwTile;
    const int x_chunks = 3 * kCwHalo * cx4, all_chunks = x_chunks + kCwTile * cg4;
    bool ok = true;

    // 16-byte cp.async of one tile's raw rows; rows outside [0, n) are skipped (their lanes are masked later)
    auto fetch = [&](int64_t tile, int b) {
        const int64_t v0 = tile * kCwTile;
        const unsigned dst0 = raw0 + b * raw_bytes;
        for (int e = tid; e < all_chunks; e += kCwThreads) {
            if (e < x_chunks) {
                const int p = e / (kCwHalo * cx4), r = (e / cx4) % kCwHalo, c = e % cx4;
                const int64_t row = v0 + (int64_t)(p - 1) * Y - 1 + r;
                if (row >= 0 && row < n) {
                    const unsigned d = dst0 + (unsigned)((p * kCwHalo + r) * ldx + 4 * c) * 4u;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(x + row * cx + 4 * c) : "memory");
                }
            } else {
                const int e2 = e - x_chunks, r = e2 / cg4, c = e2 % cg4;
                const int64_t row = v0 + r;
                if (row < n) {
                    const unsigned d = dst0 + (unsigned)(3 * kCwHalo * ldx + r * ldg_ + 4 * c) * 4u;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(g + row * cg + 4 * c) : "memory");
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    int it = 0;
    if ((int64_t)blockIdx.x < tiles) fetch(blockIdx.x, 0);
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        const int b = it & 1;
        const int64_t next = tile + gridDim.x;
        if (next < tiles) {
            fetch(next, b ^ 1);                                          // raw[b^1] was last read two barriers ago
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();                                                 // every thread's chunks of this tile landed
        if (it >= 1) ok = cw_mbar_wait(cw_smem_u32(sBar), (it - 1) & 1) && ok;     // MMAs of the previous tile left the operands
        cw_after_sync();
        // this lane's voxel and the taps that stay inside the image
        const int64_t v = tile * kCwTile + lane;
        const bool in = v < n;
        const int yy = (int)(v % Y), xx = (int)((v / Y) % X);
        unsigned tapmask = 0;
        if (in) {
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const int dx = t / 3 - 1, dy = t % 3 - 1;
                if (xx + dx >= 0 && xx + dx < X && yy + dy >= 0 && yy + dy < Y) tapmask |= 1u << t;
            }
        }
        const unsigned rawb = raw0 + b * raw_bytes;
        for (int task = warp; task < n_tasks; task += kCwThreads / 32) {
            float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
            unsigned dst;
            int row0;
            if (task < 9 * cx4) {
                const int t = task / cx4, c = task - t * cx4;
                const int p = t / 3, dy = t % 3 - 1;
                if ((tapmask >> t) & 1u) {
                    const unsigned src = rawb + (unsigned)((p * kCwHalo + lane + dy + 1) * ldx + 4 * c) * 4u;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                 : "=f"(val.x), "=f"(val.y), "=f"(val.z), "=f"(val.w) : "r"(src));
                }
                dst = base + (t >> 1) * kCwATile;
                row0 = (t & 1) * 64 + 4 * c;
            } else {
                const int c = task - 9 * cx4;
                if (in) {
                    const unsigned src = rawb + (unsigned)(3 * kCwHalo * ldx + lane * ldg_ + 4 * c) * 4u;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                 : "=f"(val.x), "=f"(val.y), "=f"(val.z), "=f"(val.w) : "r"(src));
                }
                dst = base + kCwGroups * kCwATile;
                row0 = 4 * c;
            }
            cw_sts(dst + cw_off(row0 + 0, lane), val.x);
            cw_sts(dst + cw_off(row0 + 1, lane), val.y);
            cw_sts(dst + cw_off(row0 + 2, lane), val.z);
            cw_sts(dst + cw_off(row0 + 3, lane), val.w);
        }
        cw_before_sync();
        cw_fence_async();
        __syncthreads();
        if (tid == 0) {
            cw_after_sync();
            const unsigned bt = base + kCwGroups * kCwATile;
#pragma unroll
            for (int grp = 0; grp < kCwGroups; ++grp) {
                const unsigned at = base + grp * kCwATile;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    cw_mma(tmem + grp * 64, cw_desc(at) + 2 * k, cw_desc(bt) + 2 * k, idesc, (it > 0 || k > 0) ? 1u : 0u);
            }
            cw_commit(cw_smem_u32(sBar));
        }
    }
    // the commit of the last tile covers every MMA issued before it
    if (it > 0) ok = cw_mbar_wait(cw_smem_u32(sBar), (it - 1) & 1) && ok;
    cw_after_sync();
    if (it > 0) {
        float* out = partial + (int64_t)blockIdx.x * kCwPartial;
        const int q = warp & 3;                                    // a warp reads TMEM lanes 32 q .. 32 q + 31
        for (int grp = warp >> 2; grp < kCwGroups; grp += kCwThreads / 128) {
            const unsigned taddr = tmem + grp * 64 + ((unsigned)(q * 32) << 16);
            float* dst = out + ((int64_t)grp * 128 + q * 32 + lane) * 64;
#pragma unroll
            for (int part = 0; part < 4; ++part) {
                float acc[16];
                cw_tmem_ld16(taddr + part * 16, acc);
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    *reinterpret_cast<float4*>(dst + part * 16 + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
            }
        }
    }
    if (!ok && status != nullptr) atomicExch(status, 1);
    cw_before_sync();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((unsigned)kCwTmemCols)
                     : "memory");
    }
}

// dw [cg, cx, 3, 3] (+)= sum over CTAs of partial[cta][tap >> 1][(tap & 1) * 64 + i][o], fixed order.
__global__ void __launch_bounds__(256) k_conv_wgrad_reduce(const float* __restrict__ partial, int n_parts, int cg, int cx,
                                                          float* __restrict__ dw, int accumulate) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;          // e = (grp * 128 + m) * 64 + o
    if (e >= kCwPartial) return;
    const int o = e & 63, m = (e >> 6) & 127, grp = e >> 13;
    const int tap = 2 * grp + (m >> 6), i = m & 63;
    if (tap > 8 || o >= cg || i >= cx) return;
    float s = 0.f;
    for (int p = 0; p < n_parts; ++p) s += partial[(int64_t)p * kCwPartial + e];
    float* d = dw + ((int64_t)o * cx + i) * 9 + tap;               // tap = 3 kx + ky
    *d = accumulate ? *d + s : s;
}

}  // namespace qb

using namespace qb;

extern "C" int64_t qbold_conv_wgrad_workspace_floats(void) { return (int64_t)sm_count() * kCwPartial; }

extern "C" int qbold_conv_wgrad(const float* g, int32_t cg, const float* x, int32_t cx, int64_t n_images, int32_t nx,
                                int32_t ny, float* dw, int32_t accumulate, float* workspace, int32_t* status,
                                void* stream) {
    if (cg < 4 || cg > 64 || (cg & 3) || cx < 4 || cx > 64 || (cx & 3) || n_images < 0 || nx < 1 || ny < 1)
        return fail(QBOLD_EUNSUPPORTED, "qbold_conv_wgrad: channels must be multiples of 4 in [4, 64] (got %d, %d)", cg, cx);
    if (!g || !x || !dw || !workspace) return fail(QBOLD_EINVAL, "qbold_conv_wgrad: null pointer");
    if ((reinterpret_cast<uintptr_t>(g) & 15) || (reinterpret_cast<uintptr_t>(x) & 15) ||
        (reinterpret_cast<uintptr_t>(workspace) & 15))
        return fail(QBOLD_EINVAL, "qbold_conv_wgrad: g, x, workspace must be 16-byte aligned");
    const int64_t n = n_images * nx * ny;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t grid = 0;
    if (n > 0) {
        const size_t smem = 1024 + (size_t)kCwBuf + 2 * (size_t)kCwRawMax + 64;
        int rc = cuda_check(cudaFuncSetAttribute(k_conv_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                            "cudaFuncSetAttribute(k_conv_wgrad_tc)");
        if (rc) return rc;
        const int64_t tiles = (n + kCwTile - 1) / kCwTile;
        grid = (int64_t)sm_count();
        if (tiles < grid) grid = tiles;
        k_conv_wgrad_tc<<<(unsigned)grid, kCwThreads, smem, st>>>(g, cg, x, cx, n, nx, ny, workspace, status);
        rc = after_launch("k_conv_wgrad_tc");
        if (rc) return rc;
    }
    k_conv_wgrad_reduce<<<(kCwPartial + 255) / 256, 256, 0, st>>>(workspace, (int)grid, cg, cx, dw, accumulate);
    return after_launch("k_conv_wgrad_reduce");
}
