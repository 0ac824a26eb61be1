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
// A = x line [M = (tap ky, 32 channels i), K = voxel], B = g line [N = channel o, K = voxel], both MN-major in the
// SWIZZLE_128B_BASE32B layout (TMA: 128B_ATOM_32B), 32-channel x 4-voxel swizzle atoms.  The three ky taps are the SAME
// shared-memory rows: the descriptor's leading-dimension stride between the four M atoms is ONE ROW (128 bytes), so atom
// s reads the line moved by s voxels (the swizzle is a function of the absolute shared-memory address, which TMA and
// tcgen05 share; the fourth atom is unused).  The three kx taps are the lines above / at / below, which a CTA that
// walks down consecutive lines keeps in a ring: every x line is fetched from L2 ONCE per CTA and every g line once.
//   per image line: 2 TMA boxes of x (Y + 2 voxels incl. the zero-filled halo, two channel halves) and 2 of g;
//   Y / 8 k-steps x 3 (kx) x 2 (channel half) tcgen05.mma kind::tf32 M=128 N=64 K=8 into six 64-column TMEM accumulators
//   that live for the whole kernel (48 clk each: bound by the 128 B/clk shared-memory operand fetch, see
//   tools/micro/umma_rate.cu); warp 0 produces (TMA), warps 1 and 2 issue MMAs, full / empty mbarriers per ring slot; at the
//   end three warps dump TMEM to a per-CTA partial, summed in fixed order by k_conv_wgrad_tma_reduce.
#include <cstdlib>

#include "tma_umma.cuh"

namespace qb {

namespace {

constexpr int kCtThreads = 128;
constexpr int kCtAcc = 6;                      // accumulators [line offset kx][channel half]
constexpr int kCtPartial = kCtAcc * 128 * 64;  // floats per CTA
constexpr int kCtTmemCols = 512;
constexpr int kCtXSlots = 6, kCtGSlots = 3;    // rings of x lines / g lines

struct WgradGeom {
    int X, Y, by, ny_tiles;
    int rows_a;              // shared-memory rows reserved per x box: by + 2 rounded up to 8
    long long n_images;
    long long lines;         // output lines: ny_tiles * n_images * X
    long long lines_per_cta;
    int debug;               // QBOLD_CONV_WGRAD_DEBUG: 1 = no MMAs (loads only), 2 = no loads (MMAs on stale data)
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
    const unsigned half_x = (unsigned)geo.rows_a * 128u, slot_x = 2u * half_x;      // one x line: two channel halves
    const unsigned half_g = (unsigned)geo.by * 128u, slot_g = 2u * half_g;
    const unsigned g_base = base + kCtXSlots * slot_x;
    const unsigned bar0 = g_base + kCtGSlots * slot_g;
    // barriers: x full / empty [kCtXSlots], g full / empty [kCtGSlots], done; then the TMEM base address
    constexpr int kBarsN = 2 * kCtXSlots + 2 * kCtGSlots + 1;
    unsigned* sTmem = reinterpret_cast<unsigned*>(smem_raw + (bar0 - raw) + 8 * kBarsN);
    const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    auto xfull = [&](int s) { return bar0 + 8u * (unsigned)s; };
    auto xempty = [&](int s) { return bar0 + 8u * (unsigned)(kCtXSlots + s); };
    auto gfull = [&](int s) { return bar0 + 8u * (unsigned)(2 * kCtXSlots + s); };
    auto gempty = [&](int s) { return bar0 + 8u * (unsigned)(2 * kCtXSlots + kCtGSlots + s); };
    const unsigned done_bar = bar0 + 8u * (unsigned)(kBarsN - 1);

    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ct_smem_u32(sTmem)),
                     "r"((unsigned)kCtTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int b = 0; b < kCtXSlots; ++b) ct_mbar_init(xfull(b), 1), ct_mbar_init(xempty(b), 2);   // both issuers release
        for (int b = 0; b < kCtGSlots; ++b) ct_mbar_init(gfull(b), 1), ct_mbar_init(gempty(b), 2);
        ct_mbar_init(done_bar, 2);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_g) : "memory");
    }
    ct_before_sync();
    __syncthreads();
    ct_after_sync();
    const unsigned tmem = *sTmem;
    bool ok = true;
    // This CTA's output lines [l0, l1): line l = ((y segment * n_images + image) * X + x).  In the padded numbering
    // p = l + 2 (l / X) + 1 every image has a zero line above and below (x = -1, X: zero-filled by TMA), so the CTA's
    // lines, pads included, are the consecutive p in [pc0, pc1] and the x lines it needs are [pc0 - 1, pc1 + 1].
    const long long l0 = (long long)blockIdx.x * geo.lines_per_cta;
    const long long l1 = l0 + geo.lines_per_cta < geo.lines ? l0 + geo.lines_per_cta : geo.lines;
    const bool active = l0 < l1;
    const long long pc0 = active ? l0 + 2 * (l0 / geo.X) + 1 : 0, pc1 = active ? (l1 - 1) + 2 * ((l1 - 1) / geo.X) + 1 : -1;
    const int xp = geo.X + 2;
    auto coords = [&](long long p, int& xr, int& img, int& y0) {       // padded line -> TMA coordinates
        xr = (int)(p % xp) - 1;
        const long long blk = p / xp;
        img = (int)(blk % geo.n_images);
        y0 = (int)(blk / geo.n_images) * geo.by;
    };

    if (warp == 0 && lane == 0 && active) {
        // ===== TMA producer: x lines in padded order (two ahead of the line being multiplied), g lines of the real ones
        const unsigned tx_x = 2u * (unsigned)(geo.by + 2) * 128u, tx_g = 2u * half_g;
        long long jx = 0, jg = 0;                                       // lines issued so far into each ring
        for (long long pc = pc0; pc <= pc1; ++pc) {
            for (; jx <= pc - pc0 + 2; ++jx) {                          // x lines pc0 - 1 + jx, up to pc + 1
                const int s = (int)(jx % kCtXSlots);
                ok = ct_mbar_wait_relaxed(xempty(s), (unsigned)((jx / kCtXSlots) & 1) ^ 1u, 200u) && ok;
                int xr, img, y0;
                coords(pc0 - 1 + jx, xr, img, y0);
                if (geo.debug == 2) {
                    ct_mbar_arrive(xfull(s));
                    continue;
                }
                ct_mbar_expect_tx(xfull(s), tx_x);
                ct_tma_4d(base + (unsigned)s * slot_x, &tm_x, 0, y0 - 1, xr, img, xfull(s));
                ct_tma_4d(base + (unsigned)s * slot_x + half_x, &tm_x, 32, y0 - 1, xr, img, xfull(s));
            }
            int xr, img, y0;
            coords(pc, xr, img, y0);
            if (xr >= 0 && xr < geo.X) {
                const int s = (int)(jg % kCtGSlots);
                ok = ct_mbar_wait_relaxed(gempty(s), (unsigned)((jg / kCtGSlots) & 1) ^ 1u, 200u) && ok;
                if (geo.debug == 2) {
                    ct_mbar_arrive(gfull(s));
                } else {
                    ct_mbar_expect_tx(gfull(s), tx_g);
                    ct_tma_4d(g_base + (unsigned)s * slot_g, &tm_g, 0, y0, xr, img, gfull(s));
                    ct_tma_4d(g_base + (unsigned)s * slot_g + half_g, &tm_g, 32, y0, xr, img, gfull(s));
                }
                ++jg;
            }
        }
    } else if ((warp == 1 || warp == 2) && active) {
        // ===== two MMA issuer warps (whole warps, see ct_elect_one), three accumulators each: even so the issue of one
        // tcgen05.mma costs ~70 clk (descriptor registers -> uniform registers) against 48 clk of tensor-pipe time.
        // Everything but the descriptor's address field is hoisted out of the k loop.
        const int a0 = 3 * (warp - 1);
        const unsigned idesc = ct_idesc(128, 64, true, true);
        const int ksteps = geo.by >> 3;
        int sx = 0, sg = 0;                                             // ring slots of x line pc - 1 / of the next g line
        unsigned phx = 0, phg = 0;                                      // parity of the slot that x line pc + 1 / g fills
        int sx2 = 2;                                                    // ring slot of x line pc + 1
        bool first = true;
        ok = ct_mbar_wait(xfull(0), 0u) && ok;
        ok = ct_mbar_wait(xfull(1), 0u) && ok;
        int xr = (int)(pc0 % xp) - 1;
        for (long long pc = pc0; pc <= pc1; ++pc) {
            ok = ct_mbar_wait(xfull(sx2), phx) && ok;
            if (xr >= 0 && xr < geo.X) {
                ok = ct_mbar_wait(gfull(sg), phg) && ok;
                ct_after_sync();
                uint64_t da[3];
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    const int a = a0 + t;                               // accumulator = 2 * (line offset d) + channel half
                    int sl = sx + (a >> 1);
                    if (sl >= kCtXSlots) sl -= kCtXSlots;
                    // A: 32 input channels of x line d, M atoms = the same rows moved by 0, 1, 2, (3) voxels = ky
                    da[t] = ct_desc(base + (unsigned)sl * slot_x + (unsigned)(a & 1) * half_x, 128u, 512u, kLayoutSw128Base32);
                }
                // B: 8 voxels of the g line per k-step, both output halves (N atoms one half apart)
                uint64_t db = ct_desc(g_base + (unsigned)sg * slot_g, half_g, 512u, kLayoutSw128Base32);
                unsigned acc = first ? 0u : 1u;
#pragma unroll 1
                for (int j = 0; j < (geo.debug == 1 ? 0 : ksteps); ++j) {
#pragma unroll
                    for (int t = 0; t < 3; ++t) {
                        if (ct_elect_one()) ct_mma(tmem + (unsigned)(a0 + t) * 64u, da[t], db, idesc, acc);
                        da[t] += 64;                                    // 8 rows = 1024 bytes, in 16-byte units
                    }
                    db += 64;
                    acc = 1u;
                }
                first = false;
                if (ct_elect_one()) ct_commit(gempty(sg));
                if (++sg == kCtGSlots) {
                    sg = 0;
                    phg ^= 1u;
                }
            }
            if (ct_elect_one()) ct_commit(xempty(sx));                  // x line pc - 1 is not needed again
            if (++sx == kCtXSlots) sx = 0;
            if (++sx2 == kCtXSlots) {
                sx2 = 0;
                phx ^= 1u;
            }
            if (++xr > geo.X) xr = -1;
        }
        if (ct_elect_one()) ct_commit(done_bar);                        // covers every MMA issued by this warp's elected lane
    }
    __syncwarp();
    // ===== epilogue: warp w owns TMEM lanes 32 w .. 32 w + 31 = tap ky = w (w = 3: the unused fourth shift)
    if (active) {
        ok = ct_mbar_wait_relaxed(done_bar, 0u, 2000u) && ok;
        ct_after_sync();
        float* out = partial + (long long)blockIdx.x * kCtPartial;
        if (warp < 3) {
            for (int a = 0; a < kCtAcc; ++a) {
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
    }
    if (!ok && status != nullptr) atomicExch(status, 1);
    ct_before_sync();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((unsigned)kCtTmemCols)
                     : "memory");
    }
}

// dw [cg, cx, 3, 3] (+)= sum over CTAs of partial[cta][2 kx + i / 32][32 ky + i % 32][o], fixed order.
__global__ void __launch_bounds__(256) k_conv_wgrad_tma_reduce(const float* __restrict__ partial, int n_parts, int cg,
                                                              int cx, float* __restrict__ dw, int accumulate) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;          // e = ((kx * 3 + ky) * 64 + i) * 64 + o
    if (e >= 9 * 64 * 64) return;
    const int o = e & 63, i = (e >> 6) & 63, tap = e >> 12;
    const int kx = tap / 3, ky = tap % 3;
    if (o >= cg || i >= cx) return;
    const long long src = ((long long)(2 * kx + (i >> 5)) * 128 + ky * 32 + (i & 31)) * 64 + o;
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
        geo.by = ny >= 64 ? 64 : ((ny + 7) & ~7);
        geo.ny_tiles = (ny + geo.by - 1) / geo.by;
        geo.rows_a = (geo.by + 2 + 7) & ~7;
        geo.n_images = n_images;
        geo.lines = (long long)geo.ny_tiles * n_images * nx;
        grid = (int64_t)sm_count();
        if (geo.lines < grid) grid = geo.lines;
        geo.lines_per_cta = (geo.lines + grid - 1) / grid;
        const char* dbg = getenv("QBOLD_CONV_WGRAD_DEBUG");
        geo.debug = dbg ? atoi(dbg) : 0;
        grid = (geo.lines + geo.lines_per_cta - 1) / geo.lines_per_cta;
        const size_t smem = 1024 + (size_t)kCtXSlots * 2 * geo.rows_a * 128 + (size_t)kCtGSlots * 2 * geo.by * 128 +
                            8 * (2 * kCtXSlots + 2 * kCtGSlots + 1) + 16;
        CUtensorMap tm_x, tm_g;
        int rc = activation_map(&tm_x, x, cx, n_images, nx, ny, geo.by + 2, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
        if (rc) return rc;
        rc = activation_map(&tm_g, g, cg, n_images, nx, ny, geo.by, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
        if (rc) return rc;
        rc = cuda_check(cudaFuncSetAttribute(k_conv_wgrad_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                        "cudaFuncSetAttribute(k_conv_wgrad_tma)");
        if (rc) return rc;
        k_conv_wgrad_tma<<<(unsigned)grid, kCtThreads, smem, st>>>(tm_x, tm_g, geo, workspace, status);
        rc = after_launch("k_conv_wgrad_tma");
        if (rc) return rc;
    }
    k_conv_wgrad_tma_reduce<<<(9 * 64 * 64 + 255) / 256, 256, 0, st>>>(workspace, (int)grid, cg, cx, dw, accumulate);
    return after_launch("k_conv_wgrad_tma_reduce");
}
