// The encoder's 3x3x1 convolutions (reference create_block, model.py:152,156; padding 'same') as TMA-fed tcgen05 GEMMs.
//
// Activations are kept z-outer, [B*Z, X, Y, C] float32 (qbold_vi_b200/encoder.py), C <= 64 a multiple of 4.  A 4-D TMA
// tensor map (C, Y, X, B*Z) with SWIZZLE_128B and a box of 32 channels x (rows of one image line) delivers any window
// of an image line straight into the canonical tcgen05 shared-memory layout; coordinates outside the image (the
// 'same' padding, the line above the first and below the last, channels 60..63) are zero-filled by the TMA unit, so
// there is no im2col, no halo logic and no register staging anywhere.
//
// Weight gradient (this file, first kernel): a GEMM whose contraction runs over the VOXELS,
//     dW[o, i, kx, ky] = sum_v g[v, o] * x[v + (kx-1, ky-1), i].
// With the voxel index as K, the activations as they lie in memory ARE the MN-major operands (channels contiguous):
// A = x window [M = (line offset dx, channel i), K = voxel], B = g tile [N = channel o, K = voxel], both MN-major in
// the SWIZZLE_128B_BASE32B layout (TMA: 128B_ATOM_32B), 32-channel x 4-voxel swizzle atoms.  The three ky taps of one
// line offset are the SAME shared-memory rows read through descriptors whose start address is moved by 0 / 1 / 2 rows
// (128 bytes) -- the swizzle is a function of the absolute shared-memory address, which TMA and tcgen05 share -- so x
// is fetched 3 x (BY + 2) / BY times instead of 9 times.
//   per tile (one image line segment of BY voxels): 6 TMA boxes of x (3 lines x 2 channel halves, BY + 2 voxels),
//   2 of g; 3 (ky) x 2 x BY/8 tcgen05.mma kind::tf32 M=128 N=64 K=8:
//     MMA 1: M rows = (dx=0, i), (dx=1, i)      -> accumulator [ky][0]
//     MMA 2: M rows = (dx=2, i), (junk)         -> accumulator [ky][1]   (upper 64 lanes unused)
//   six 64-column TMEM accumulators live for the whole kernel; warp 0 produces (TMA), warp 1 issues MMAs, full / empty
//   mbarriers over a 5-stage ring; at the end the four warps dump TMEM to a per-CTA partial, summed in fixed order.
#include "tma_umma.cuh"

namespace qb {

namespace {

constexpr int kCtThreads = 128;
constexpr int kCtAcc = 6;                      // accumulators [ky][second MMA]
constexpr int kCtPartial = kCtAcc * 128 * 64;  // floats per CTA
constexpr int kCtTmemCols = 512;
constexpr int kCtMaxStages = 6;
constexpr size_t kCtSmemBudget = 220 * 1024;

struct WgradGeom {
    int X, Y, by, ny_tiles, stages;
    int rows_a;          // shared-memory rows reserved per x box: by + 2 rounded up to 8
    long long tiles;     // n_images * X * ny_tiles
};

}  // namespace

// partial [gridDim.x][6][128][64]
__global__ void __launch_bounds__(kCtThreads, 1) k_conv_wgrad_tma(const __grid_constant__ CUtensorMap tm_x,
                                                                  const __grid_constant__ CUtensorMap tm_g,
                                                                  const WgradGeom geo, float* __restrict__ partial,
                                                                  int* __restrict__ status) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const unsigned raw = ct_smem_u32(smem_raw);
    const unsigned base = (raw + 1023u) & ~1023u;
    const unsigned lbo_a = (unsigned)geo.rows_a * 128u, lbo_b = (unsigned)geo.by * 128u;
    const unsigned stage_bytes = 6u * lbo_a + 2u * lbo_b;
    const unsigned tail = base + (unsigned)geo.stages * stage_bytes + 2u * lbo_a;   // slack: the junk atoms of MMA 2
    // barriers: full[stages], empty[stages], done; then the TMEM base address
    const unsigned bar0 = (tail + 15u) & ~15u;
    unsigned* sTmem = reinterpret_cast<unsigned*>(smem_raw + (bar0 - raw) + 8 * (2 * kCtMaxStages + 1));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    auto full_bar = [&](int s) { return bar0 + 8u * (unsigned)s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (unsigned)(kCtMaxStages + s); };
    const unsigned done_bar = bar0 + 8u * (unsigned)(2 * kCtMaxStages);

    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ct_smem_u32(sTmem)),
                     "r"((unsigned)kCtTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int s = 0; s < geo.stages; ++s) {
            ct_mbar_init(full_bar(s), 1);
            ct_mbar_init(empty_bar(s), 1);
        }
        ct_mbar_init(done_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_g) : "memory");
    }
    ct_before_sync();
    __syncthreads();
    ct_after_sync();
    const unsigned tmem = *sTmem;
    bool ok = true;
    const long long my_tiles = geo.tiles > (long long)blockIdx.x
                                   ? (geo.tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (warp == 0 && lane == 0) {
        // ===== TMA producer
        const unsigned tx = 6u * (unsigned)(geo.by + 2) * 128u + 2u * lbo_b;
        int s = 0;
        unsigned ph = 0;
        for (long long it = 0; it < my_tiles; ++it) {
            const long long tile = (long long)blockIdx.x + it * gridDim.x;
            const int yt = (int)(tile % geo.ny_tiles);
            const long long line = tile / geo.ny_tiles;
            const int xr = (int)(line % geo.X), img = (int)(line / geo.X);
            ok = ct_mbar_wait(empty_bar(s), ph ^ 1u) && ok;
            const unsigned st = base + (unsigned)s * stage_bytes;
            ct_mbar_expect_tx(full_bar(s), tx);
            const int y0 = yt * geo.by;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                ct_tma_4d(st + (unsigned)(2 * dx) * lbo_a, &tm_x, 0, y0 - 1, xr + dx - 1, img, full_bar(s));
                ct_tma_4d(st + (unsigned)(2 * dx + 1) * lbo_a, &tm_x, 32, y0 - 1, xr + dx - 1, img, full_bar(s));
            }
            ct_tma_4d(st + 6u * lbo_a, &tm_g, 0, y0, xr, img, full_bar(s));
            ct_tma_4d(st + 6u * lbo_a + lbo_b, &tm_g, 32, y0, xr, img, full_bar(s));
            if (++s == geo.stages) {
                s = 0;
                ph ^= 1u;
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ===== MMA issuer
        const unsigned idesc = ct_idesc(128, 64, true, true);
        int s = 0;
        unsigned ph = 0;
        const int ksteps = geo.by >> 3;
        for (long long it = 0; it < my_tiles; ++it) {
            ok = ct_mbar_wait(full_bar(s), ph) && ok;
            ct_after_sync();
            const unsigned st = base + (unsigned)s * stage_bytes;
            const unsigned bt = st + 6u * lbo_a;
            for (int j = 0; j < ksteps; ++j) {
                const uint64_t db = ct_desc(bt + (unsigned)j * 1024u, lbo_b, 512u, kLayoutSw128Base32);
                const unsigned acc = (it > 0 || j > 0) ? 1u : 0u;
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    const unsigned row = (unsigned)(8 * j + ky) * 128u;
                    ct_mma(tmem + (unsigned)(2 * ky) * 64u, ct_desc(st + row, lbo_a, 512u, kLayoutSw128Base32), db, idesc, acc);
                    ct_mma(tmem + (unsigned)(2 * ky + 1) * 64u,
                           ct_desc(st + 4u * lbo_a + row, lbo_a, 512u, kLayoutSw128Base32), db, idesc, acc);
                }
            }
            ct_commit(empty_bar(s));                         // frees the stage once these MMAs have read it
            if (++s == geo.stages) {
                s = 0;
                ph ^= 1u;
            }
        }
        ct_commit(done_bar);                                  // covers every MMA issued by this thread
    }
    __syncwarp();
    // ===== epilogue: all four warps, warp w owns TMEM lanes 32 w .. 32 w + 31
    if (my_tiles > 0) {
        ok = ct_mbar_wait(done_bar, 0u) && ok;
        ct_after_sync();
        float* out = partial + (long long)blockIdx.x * kCtPartial;
        for (int a = 0; a < kCtAcc; ++a) {
            if ((a & 1) && warp >= 2) continue;              // second MMA: lanes 64..127 are junk
            const unsigned taddr = tmem + (unsigned)a * 64u + ((unsigned)(warp * 32) << 16);
            float* dst = out + ((long long)a * 128 + warp * 32 + lane) * 64;
#pragma unroll
            for (int part = 0; part < 4; ++part) {
                float acc[16];
                ct_tmem_ld16(taddr + part * 16, acc);
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    *reinterpret_cast<float4*>(dst + part * 16 + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
            }
        }
    }
    if (!ok && status != nullptr) atomicExch(status, 1);
    ct_before_sync();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((unsigned)kCtTmemCols)
                     : "memory");
    }
}

// dw [cg, cx, 3, 3] (+)= sum over CTAs of partial[cta][2 ky + (kx == 2)][(kx & 1) * 64 + i][o], fixed order.
__global__ void __launch_bounds__(256) k_conv_wgrad_tma_reduce(const float* __restrict__ partial, int n_parts, int cg,
                                                              int cx, float* __restrict__ dw, int accumulate) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;          // e = ((ky * 3 + kx) * 64 + i) * 64 + o
    if (e >= 9 * 64 * 64) return;
    const int o = e & 63, i = (e >> 6) & 63, tap = e >> 12;
    const int ky = tap / 3, kx = tap % 3;
    if (o >= cg || i >= cx) return;
    const long long src = ((long long)(2 * ky + (kx == 2)) * 128 + (kx & 1) * 64 + i) * 64 + o;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int p = 0;
    for (; p + 4 <= n_parts; p += 4) {
        s0 += partial[(long long)p * kCtPartial + src];
        s1 += partial[(long long)(p + 1) * kCtPartial + src];
        s2 += partial[(long long)(p + 2) * kCtPartial + src];
        s3 += partial[(long long)(p + 3) * kCtPartial + src];
    }
    for (; p < n_parts; ++p) s0 += partial[(long long)p * kCtPartial + src];
    const float s = (s0 + s1) + (s2 + s3);
    float* d = dw + ((long long)o * cx + i) * 9 + kx * 3 + ky;
    *d = accumulate ? *d + s : s;
}

EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// Tensor map over a z-outer activation [n_images, X, Y, C] float32: dims (C, Y, X, n_images), box 32 channels x
// box_y voxels of one image line, SWIZZLE_128B, zero fill outside.
int activation_map(CUtensorMap* tm, const float* ptr, int c, long long n_images, int X, int Y, int box_y,
                   CUtensorMapSwizzle swizzle) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return fail(QBOLD_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)Y, (cuuint64_t)X, (cuuint64_t)n_images};
    const cuuint64_t strides[3] = {(cuuint64_t)c * 4, (cuuint64_t)Y * c * 4, (cuuint64_t)X * Y * c * 4};
    const cuuint32_t box[4] = {32, (cuuint32_t)box_y, 1, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(ptr), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(QBOLD_ECUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}

int matrix_map(CUtensorMap* tm, const float* ptr, long long rows, int cols, int box_rows, CUtensorMapSwizzle swizzle) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return fail(QBOLD_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
    const cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(QBOLD_ECUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}

}  // namespace qb

using namespace qb;

extern "C" int64_t qbold_conv_wgrad_workspace_floats(void) { return (int64_t)sm_count() * kCtPartial; }

extern "C" int qbold_conv_wgrad(const float* g, int32_t cg, const float* x, int32_t cx, int64_t n_images, int32_t nx,
                                int32_t ny, float* dw, int32_t accumulate, float* workspace, int32_t* status,
                                void* stream) {
    if (cg < 4 || cg > 64 || (cg & 3) || cx < 4 || cx > 64 || (cx & 3) || n_images < 0 || nx < 1 || ny < 1)
        return fail(QBOLD_EUNSUPPORTED, "qbold_conv_wgrad: channels must be multiples of 4 in [4, 64] (got %d, %d)", cg, cx);
    if (!g || !x || !dw || !workspace) return fail(QBOLD_EINVAL, "qbold_conv_wgrad: null pointer");
    if ((reinterpret_cast<uintptr_t>(g) & 15) || (reinterpret_cast<uintptr_t>(x) & 15) ||
        (reinterpret_cast<uintptr_t>(workspace) & 15))
        return fail(QBOLD_EINVAL, "qbold_conv_wgrad: g, x, workspace must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    int64_t grid = 0;
    if (n_images > 0) {
        WgradGeom geo;
        geo.X = nx;
        geo.Y = ny;
        geo.by = ny > 16 ? 32 : (ny > 8 ? 16 : 8);
        geo.ny_tiles = (ny + geo.by - 1) / geo.by;
        geo.rows_a = (geo.by + 2 + 7) & ~7;
        geo.tiles = (long long)n_images * nx * geo.ny_tiles;
        const size_t stage_bytes = 6 * (size_t)geo.rows_a * 128 + 2 * (size_t)geo.by * 128;
        const size_t fixed = 1024 + 2 * (size_t)geo.rows_a * 128 + 16 + 8 * (2 * kCtMaxStages + 1) + 16;
        int stages = (int)((kCtSmemBudget - fixed) / stage_bytes);
        if (stages > kCtMaxStages) stages = kCtMaxStages;
        geo.stages = stages;
        const size_t smem = fixed + stages * stage_bytes;
        CUtensorMap tm_x, tm_g;
        int rc = activation_map(&tm_x, x, cx, n_images, nx, ny, geo.by + 2, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
        if (rc) return rc;
        rc = activation_map(&tm_g, g, cg, n_images, nx, ny, geo.by, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
        if (rc) return rc;
        rc = cuda_check(cudaFuncSetAttribute(k_conv_wgrad_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                        "cudaFuncSetAttribute(k_conv_wgrad_tma)");
        if (rc) return rc;
        grid = (int64_t)sm_count();
        if (geo.tiles < grid) grid = geo.tiles;
        k_conv_wgrad_tma<<<(unsigned)grid, kCtThreads, smem, st>>>(tm_x, tm_g, geo, workspace, status);
        rc = after_launch("k_conv_wgrad_tma");
        if (rc) return rc;
    }
    k_conv_wgrad_tma_reduce<<<(9 * 64 * 64 + 255) / 256, 256, 0, st>>>(workspace, (int)grid, cg, cx, dw, accumulate);
    return after_launch("k_conv_wgrad_tma_reduce");
}
