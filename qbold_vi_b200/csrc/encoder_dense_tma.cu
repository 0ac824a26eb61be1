// One per-voxel Dense layer of the encoder (reference create_layer: Conv3D 1x1x1, model.py:115-120) for the training
// passes, as a persistent TMA -> tcgen05 -> TMA pipeline:
//     y[n, n_out] = act(x[n, n_in] B + bias) (+ addend),   B = w^T (forward, w [n_out, n_in])  or  w (input gradient)
// n ~ 5e5 rows, <= 64 features: 4.3 GFLOP on 252 MB -- bound by HBM (40 us), not by the tensor cores (4 us), so the
// kernel is a copy engine with a GEMM in the middle:
//   warp 0      TMA producer: x tiles of 128 rows (two 32-column boxes, SWIZZLE_128B, K-major A operand) into a 3-stage
//               ring; the weight matrix once (forward: K-major B, SWIZZLE_128B; input gradient: the same rows read
//               MN-major, SWIZZLE_128B_BASE32B -- no transposed copy of w is ever made);
//   warp 1      one thread issues 8 (or 4) tcgen05.mma kind::tf32 M=128 N=64 K=8 per tile into one of two TMEM
//               accumulators; tcgen05.commit releases the stage and hands the accumulator to the epilogue;
//   warps 2-5   epilogue: TMEM -> registers (lane = row), bias / ReLU / addend, swizzled store into a shared-memory
//               staging tile that one thread writes back with a TMA store (rows beyond n and columns beyond n_out are
//               clipped by the TMA unit).  The addend tile (beta = 1 accumulation of the backward pass) is TMA-loaded
//               into the same staging tile one tile ahead.
#include "tma_umma.cuh"

namespace qb {

namespace {

constexpr int kDtThreads = 192;
constexpr int kDtRows = 128;                    // rows per tile
constexpr int kDtBox = kDtRows * 128;           // bytes of one 32-column box
constexpr int kDtStages = 3;
constexpr int kDtW = 2 * 64 * 128;              // weight image: two 32-column boxes of 64 rows
constexpr int kDtTmemCols = 128;

// barrier slots
enum { kFull = 0, kEmpty = kDtStages, kWFull = 2 * kDtStages, kAccFull = kWFull + 1, kAccEmpty = kAccFull + 2,
       kAddFull = kAccEmpty + 2, kBars = kAddFull + 2 };

__device__ __forceinline__ void named_bar(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

}  // namespace

__global__ void __launch_bounds__(kDtThreads, 1) k_dense_tma(const __grid_constant__ CUtensorMap tm_x,
                                                            const __grid_constant__ CUtensorMap tm_w,
                                                            const __grid_constant__ CUtensorMap tm_y,
                                                            const __grid_constant__ CUtensorMap tm_add,
                                                            const float* __restrict__ bias, int n_in, int n_out,
                                                            int transpose, int relu, int has_add, long long tiles,
                                                            int* __restrict__ status) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const unsigned raw = ct_smem_u32(smem_raw);
    const unsigned base = (raw + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - raw);
    // layout: x stages [3][2 boxes], weights [2 boxes of 64 rows], staging [2][2 boxes], barriers, tmem slot, bias
    const unsigned x_base = base, w_base = base + kDtStages * 2 * kDtBox, o_base = w_base + kDtW;
    const unsigned bar_base = o_base + 2 * 2 * kDtBox;
    unsigned* sTmem = reinterpret_cast<unsigned*>(sm + (bar_base - base) + 8 * kBars);
    float* sBias = reinterpret_cast<float*>(sm + (bar_base - base) + 8 * kBars + 16);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    auto bar = [&](int i) { return bar_base + 8u * (unsigned)i; };
    const int kblocks = n_in > 32 ? 2 : 1;

    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ct_smem_u32(sTmem)),
                     "r"((unsigned)kDtTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int s = 0; s < kDtStages; ++s) {
            ct_mbar_init(bar(kFull + s), 1);
            ct_mbar_init(bar(kEmpty + s), 1);
        }
        ct_mbar_init(bar(kWFull), 1);
        for (int a = 0; a < 2; ++a) {
            ct_mbar_init(bar(kAccFull + a), 1);
            ct_mbar_init(bar(kAccEmpty + a), 1);
            ct_mbar_init(bar(kAddFull + a), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 64) sBias[tid] = (bias != nullptr && tid < n_out) ? bias[tid] : 0.f;
    ct_before_sync();
    __syncthreads();
    ct_after_sync();
    const unsigned tmem = *sTmem;
    bool ok = true;
    const long long my_tiles = tiles > (long long)blockIdx.x ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer: the weights once, then the x tiles
            ct_mbar_expect_tx(bar(kWFull), (transpose ? 2u : (unsigned)kblocks) * 64u * 128u);
            if (!transpose) {
                // B[N = output o][K = input i]: rows o, 32 inputs per box
                for (int kb = 0; kb < kblocks; ++kb) ct_tma_2d(w_base + (unsigned)kb * 64u * 128u, &tm_w, 32 * kb, 0, bar(kWFull));
            } else {
                // w [n_in rows (K), n_out columns (N)]: MN-major B, 32 outputs per box, all K rows
                ct_tma_2d(w_base, &tm_w, 0, 0, bar(kWFull));
                ct_tma_2d(w_base + 64u * 128u, &tm_w, 32, 0, bar(kWFull));
            }
            int s = 0;
            unsigned ph = 0;
            for (long long it = 0; it < my_tiles; ++it) {
                const long long tile = (long long)blockIdx.x + it * gridDim.x;
                ok = ct_mbar_wait(bar(kEmpty + s), ph ^ 1u) && ok;
                ct_mbar_expect_tx(bar(kFull + s), (unsigned)(kblocks * kDtBox));
                const unsigned st = x_base + (unsigned)s * 2u * kDtBox;
                for (int kb = 0; kb < kblocks; ++kb)
                    ct_tma_2d(st + (unsigned)kb * kDtBox, &tm_x, 32 * kb, (int)(tile * kDtRows), bar(kFull + s));
                if (++s == kDtStages) {
                    s = 0;
                    ph ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer
            const unsigned idesc = ct_idesc(128, 64, false, transpose != 0);
            ok = ct_mbar_wait(bar(kWFull), 0u) && ok;
            int s = 0;
            unsigned ph = 0;
            for (long long it = 0; it < my_tiles; ++it) {
                const int a = (int)(it & 1);
                const unsigned aph = (unsigned)((it >> 1) & 1);
                ok = ct_mbar_wait(bar(kAccEmpty + a), aph ^ 1u) && ok;       // the epilogue drained this accumulator
                ok = ct_mbar_wait(bar(kFull + s), ph) && ok;
                ct_after_sync();
                const unsigned st = x_base + (unsigned)s * 2u * kDtBox;
                const int ksteps = (n_in + 7) >> 3;
                for (int k = 0; k < ksteps; ++k) {
                    const int kb = k >> 2, kk = k & 3;
                    const uint64_t da = ct_desc(st + (unsigned)kb * kDtBox + (unsigned)kk * 32u, 0u, 1024u, kLayoutSw128);
                    uint64_t db;
                    if (!transpose)
                        db = ct_desc(w_base + (unsigned)kb * 64u * 128u + (unsigned)kk * 32u, 0u, 1024u, kLayoutSw128);
                    else   // rows of w are K: 8 rows per k-step; N atoms (32 outputs) 64 rows apart
                        db = ct_desc(w_base + (unsigned)k * 8u * 128u, 64u * 128u, 512u, kLayoutSw128Base32);
                    ct_mma(tmem + (unsigned)a * 64u, da, db, idesc, k > 0 ? 1u : 0u);
                }
                ct_commit(bar(kEmpty + s));
                ct_commit(bar(kAccFull + a));
                if (++s == kDtStages) {
                    s = 0;
                    ph ^= 1u;
                }
            }
        }
    } else {
        // ===== epilogue warps 2..5: TMEM lane quarter = warp % 4, one row per lane
        const int q = warp & 3, row = q * 32 + lane;
        const bool leader = (warp == 2 && lane == 0);
        if (has_add && leader && my_tiles > 0) {
            ct_mbar_expect_tx(bar(kAddFull + 0), 2u * kDtBox);
            ct_tma_2d(o_base, &tm_add, 0, (int)((long long)blockIdx.x * kDtRows), bar(kAddFull + 0));
            ct_tma_2d(o_base + kDtBox, &tm_add, 32, (int)((long long)blockIdx.x * kDtRows), bar(kAddFull + 0));
        }
        for (long long it = 0; it < my_tiles; ++it) {
            const long long tile = (long long)blockIdx.x + it * gridDim.x;
            const int a = (int)(it & 1);
            const unsigned aph = (unsigned)((it >> 1) & 1);
            const unsigned stg = o_base + (unsigned)a * 2u * kDtBox;
            if (leader) {
                // the store of tile it-1 (other staging tile) may still be reading; the one of it-2 (this tile) must be done
                if (has_add) {
                    ct_store_wait_read<0>();
                    if (it + 1 < my_tiles) {
                        const unsigned nstg = o_base + (unsigned)(a ^ 1) * 2u * kDtBox;
                        const int r0 = (int)((tile + gridDim.x) * kDtRows);
                        ct_mbar_expect_tx(bar(kAddFull + (a ^ 1)), 2u * kDtBox);
                        ct_tma_2d(nstg, &tm_add, 0, r0, bar(kAddFull + (a ^ 1)));
                        ct_tma_2d(nstg + kDtBox, &tm_add, 32, r0, bar(kAddFull + (a ^ 1)));
                    }
                } else {
                    ct_store_wait_read<1>();
                }
            }
            named_bar(1, 128);                                            // staging tile `a` is free (or being filled)
            ok = ct_mbar_wait(bar(kAccFull + a), aph) && ok;
            if (has_add) ok = ct_mbar_wait(bar(kAddFull + a), aph) && ok;
            ct_after_sync();
            const unsigned taddr = tmem + (unsigned)a * 64u + ((unsigned)(q * 32) << 16);
#pragma unroll
            for (int part = 0; part < 4; ++part) {
                float v[16];
                ct_tmem_ld16(taddr + part * 16, v);
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    const int c = part * 16 + j;                          // column of v[j]
                    // staging tile: box c / 32, row `row`, 16-byte chunk (c % 32) / 4 swizzled with the row (SWIZZLE_128B)
                    const unsigned addr = stg + (unsigned)(c >> 5) * kDtBox + (unsigned)row * 128u +
                                          ((((unsigned)(c & 31) >> 2) ^ ((unsigned)row & 7u)) << 4);
                    float4 o = make_float4(v[j] + sBias[c], v[j + 1] + sBias[c + 1], v[j + 2] + sBias[c + 2],
                                           v[j + 3] + sBias[c + 3]);
                    if (relu) {
                        o.x = fmaxf(o.x, 0.f);
                        o.y = fmaxf(o.y, 0.f);
                        o.z = fmaxf(o.z, 0.f);
                        o.w = fmaxf(o.w, 0.f);
                    }
                    if (has_add) {
                        float4 t;
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                     : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "r"(addr));
                        o.x += t.x;
                        o.y += t.y;
                        o.z += t.z;
                        o.w += t.w;
                    }
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(o.x), "f"(o.y), "f"(o.z),
                                 "f"(o.w) : "memory");
                }
            }
            ct_fence_async();                                             // generic-proxy writes -> visible to the TMA store
            ct_before_sync();
            named_bar(2, 128);
            if (leader) {
                ct_mbar_arrive(bar(kAccEmpty + a));                       // accumulator `a` may be overwritten
                ct_tma_store_2d(&tm_y, 0, (int)(tile * kDtRows), stg);
                if (n_out > 32) ct_tma_store_2d(&tm_y, 32, (int)(tile * kDtRows), stg + kDtBox);
                ct_store_commit();
            }
        }
        if (leader) ct_store_wait_read<0>();
    }
    if (!ok && status != nullptr) atomicExch(status, 1);
    ct_before_sync();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((unsigned)kDtTmemCols)
                     : "memory");
    }
}

// ---- weight / bias gradient of a Dense layer ---------------------------------------------------------------------------
//     dW[o, i] = sum_v g[v, o] x[v, i],   db[o] = sum_v g[v, o]
// The contraction runs over the voxels, so g and x as they lie in memory are the MN-major operands (see
// csrc/encoder_conv_tma.cu): per 128-voxel tile two boxes of g (A: M = 64 outputs, upper half of M = 128 unused) and two
// of x (B: N = 64 inputs), 16 k-steps of tcgen05.mma kind::tf32 into ONE accumulator that lives for the whole kernel,
// plus a second N = 16 MMA per k-step against a tile of ones for the bias gradient.  HBM-bound (252 MB per layer).
namespace {
constexpr int kWtThreads = 128;
constexpr int kWtStages = 3;
constexpr int kWtStage = 4 * kDtBox;             // g lo, g hi, x lo, x hi
constexpr int kWtPartial = 64 * 64 + 64;         // floats per CTA: dW then db
}  // namespace

__global__ void __launch_bounds__(kWtThreads, 1) k_dense_wgrad_tma(const __grid_constant__ CUtensorMap tm_g,
                                                                  const __grid_constant__ CUtensorMap tm_x, int n_in,
                                                                  long long tiles, float* __restrict__ partial,
                                                                  int* __restrict__ status) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const unsigned raw = ct_smem_u32(smem_raw);
    const unsigned base = (raw + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - raw);
    const unsigned ones_base = base + kWtStages * kWtStage;                // 1 KB of 1.0f
    const unsigned bar_base = ones_base + 1024;
    unsigned* sTmem = reinterpret_cast<unsigned*>(sm + (bar_base - base) + 8 * (2 * kWtStages + 1));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    auto full_bar = [&](int s) { return bar_base + 8u * (unsigned)s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (unsigned)(kWtStages + s); };
    const unsigned done_bar = bar_base + 8u * (unsigned)(2 * kWtStages);
    const int xboxes = n_in > 32 ? 2 : 1;

    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ct_smem_u32(sTmem)),
                     "r"(128u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int s = 0; s < kWtStages; ++s) {
            ct_mbar_init(full_bar(s), 1);
            ct_mbar_init(empty_bar(s), 1);
        }
        ct_mbar_init(done_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 256; i += kWtThreads) reinterpret_cast<float*>(sm + (ones_base - base))[i] = 1.0f;
    ct_fence_async();
    ct_before_sync();
    __syncthreads();
    ct_after_sync();
    const unsigned tmem = *sTmem;
    bool ok = true;
    const long long my_tiles = tiles > (long long)blockIdx.x ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (warp == 0 && lane == 0) {
        int s = 0;
        unsigned ph = 0;
        for (long long it = 0; it < my_tiles; ++it) {
            const int r0 = (int)(((long long)blockIdx.x + it * gridDim.x) * kDtRows);
            ok = ct_mbar_wait(empty_bar(s), ph ^ 1u) && ok;
            const unsigned st = base + (unsigned)s * kWtStage;
            ct_mbar_expect_tx(full_bar(s), (unsigned)((2 + xboxes) * kDtBox));
            ct_tma_2d(st, &tm_g, 0, r0, full_bar(s));
            ct_tma_2d(st + kDtBox, &tm_g, 32, r0, full_bar(s));
            ct_tma_2d(st + 2 * kDtBox, &tm_x, 0, r0, full_bar(s));
            if (xboxes == 2) ct_tma_2d(st + 3 * kDtBox, &tm_x, 32, r0, full_bar(s));
            if (++s == kWtStages) {
                s = 0;
                ph ^= 1u;
            }
        }
    } else if (warp == 1 && lane == 0) {
        const unsigned idesc = ct_idesc(128, 64, true, true), idesc1 = ct_idesc(128, 16, true, true);
        const uint64_t d_ones = ct_desc(ones_base, 512u, 512u, kLayoutSw128Base32);
        int s = 0;
        unsigned ph = 0;
        for (long long it = 0; it < my_tiles; ++it) {
            ok = ct_mbar_wait(full_bar(s), ph) && ok;
            ct_after_sync();
            const unsigned st = base + (unsigned)s * kWtStage;
            for (int k = 0; k < kDtRows / 8; ++k) {
                const unsigned acc = (it > 0 || k > 0) ? 1u : 0u;
                const uint64_t da = ct_desc(st + (unsigned)k * 1024u, kDtBox, 512u, kLayoutSw128Base32);
                const uint64_t db = ct_desc(st + 2u * kDtBox + (unsigned)k * 1024u, kDtBox, 512u, kLayoutSw128Base32);
                ct_mma(tmem, da, db, idesc, acc);
                ct_mma(tmem + 64u, da, d_ones, idesc1, acc);
            }
            ct_commit(empty_bar(s));
            if (++s == kWtStages) {
                s = 0;
                ph ^= 1u;
            }
        }
        ct_commit(done_bar);
    }
    __syncwarp();
    if (my_tiles > 0 && warp < 2) {                                        // TMEM lanes 0..63 = outputs o
        ok = ct_mbar_wait(done_bar, 0u) && ok;
        ct_after_sync();
        float* out = partial + (long long)blockIdx.x * kWtPartial;
        const int o = warp * 32 + lane;
        const unsigned taddr = tmem + ((unsigned)(warp * 32) << 16);
#pragma unroll
        for (int part = 0; part < 4; ++part) {
            float acc[16];
            ct_tmem_ld16(taddr + part * 16, acc);
#pragma unroll
            for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<float4*>(out + o * 64 + part * 16 + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
        }
        float one[16];
        ct_tmem_ld16(taddr + 64, one);
        out[64 * 64 + o] = one[0];
    }
    if (!ok && status != nullptr) atomicExch(status, 1);
    ct_before_sync();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
    }
}

// dw [n_out, n_in] (+)= sum over CTAs of partial[cta][o][i]; db [n_out] (+)= partial[cta][4096 + o]; fixed order.
__global__ void __launch_bounds__(256) k_dense_wgrad_tma_reduce(const float* __restrict__ partial, int n_parts, int n_out,
                                                               int n_in, float* __restrict__ dw, float* __restrict__ db,
                                                               int accumulate) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= kWtPartial) return;
    const bool is_b = e >= 64 * 64;
    const int o = is_b ? e - 64 * 64 : e >> 6, i = e & 63;
    if (o >= n_out || (!is_b && i >= n_in) || (is_b && db == nullptr)) return;
    float s0 = 0.f, s1 = 0.f;
    int p = 0;
    for (; p + 2 <= n_parts; p += 2) {
        s0 += partial[(long long)p * kWtPartial + e];
        s1 += partial[(long long)(p + 1) * kWtPartial + e];
    }
    if (p < n_parts) s0 += partial[(long long)p * kWtPartial + e];
    const float s = s0 + s1;
    float* d = is_b ? db + o : dw + (long long)o * n_in + i;
    *d = accumulate ? *d + s : s;
}

}  // namespace qb

using namespace qb;

extern "C" int qbold_dense_tma(const float* x, const float* w, const float* bias, const float* addend, int32_t n_in,
                               int32_t n_out, int32_t transpose, int32_t relu, int64_t n, float* y, int32_t* status,
                               void* stream) {
    if (n_in < 4 || n_in > 64 || (n_in & 3) || n_out < 4 || n_out > 64 || (n_out & 3))
        return fail(QBOLD_EUNSUPPORTED, "qbold_dense_tma: n_in, n_out must be multiples of 4 in [4, 64] (got %d, %d)", n_in,
                    n_out);
    if (n < 0 || (n > 0 && (!x || !w || !y))) return fail(QBOLD_EINVAL, "qbold_dense_tma: bad argument");
    if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(w) & 15) ||
        (reinterpret_cast<uintptr_t>(y) & 15) || (reinterpret_cast<uintptr_t>(addend) & 15))
        return fail(QBOLD_EINVAL, "qbold_dense_tma: x, w, y, addend must be 16-byte aligned");
    if (n == 0) return QBOLD_OK;
    CUtensorMap tm_x, tm_w, tm_y, tm_add;
    int rc = matrix_map(&tm_x, x, n, n_in, kDtRows, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    // forward: w [n_out, n_in], boxes of 32 inputs x 64 output rows; input gradient: w [n_in, n_out] read MN-major,
    // boxes of 32 outputs x 64 input rows
    rc = transpose ? matrix_map(&tm_w, w, n_in, n_out, 64, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)
                   : matrix_map(&tm_w, w, n_out, n_in, 64, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = matrix_map(&tm_y, y, n, n_out, kDtRows, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = matrix_map(&tm_add, addend ? addend : y, n, n_out, kDtRows, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    const size_t smem = 1024 + (size_t)kDtStages * 2 * kDtBox + kDtW + 2 * 2 * kDtBox + 8 * kBars + 16 + 64 * 4;
    rc = cuda_check(cudaFuncSetAttribute(k_dense_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                    "cudaFuncSetAttribute(k_dense_tma)");
    if (rc) return rc;
    const long long tiles = (n + kDtRows - 1) / kDtRows;
    long long grid = sm_count();
    if (tiles < grid) grid = tiles;
    k_dense_tma<<<(unsigned)grid, kDtThreads, smem, (cudaStream_t)stream>>>(tm_x, tm_w, tm_y, tm_add, bias, n_in, n_out,
                                                                           transpose, relu, addend != nullptr, tiles,
                                                                           status);
    return after_launch("k_dense_tma");
}

extern "C" int64_t qbold_dense_wgrad_tma_workspace_floats(void) { return (int64_t)sm_count() * kWtPartial; }

extern "C" int qbold_dense_wgrad_tma(const float* g, int32_t n_out, const float* x, int32_t n_in, int64_t n, float* dw,
                                     float* db, int32_t accumulate, float* workspace, int32_t* status, void* stream) {
    if (n_in < 4 || n_in > 64 || (n_in & 3) || n_out < 4 || n_out > 64 || (n_out & 3))
        return fail(QBOLD_EUNSUPPORTED, "qbold_dense_wgrad_tma: n_in, n_out must be multiples of 4 in [4, 64] (got %d, %d)",
                    n_in, n_out);
    if (n < 0 || !g || !x || !dw || !workspace) return fail(QBOLD_EINVAL, "qbold_dense_wgrad_tma: bad argument");
    if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(g) & 15))
        return fail(QBOLD_EINVAL, "qbold_dense_wgrad_tma: g, x must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    long long grid = 0;
    if (n > 0) {
        CUtensorMap tm_g, tm_x;
        int rc = matrix_map(&tm_g, g, n, n_out, kDtRows, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
        if (rc) return rc;
        rc = matrix_map(&tm_x, x, n, n_in, kDtRows, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
        if (rc) return rc;
        const size_t smem = 1024 + (size_t)kWtStages * kWtStage + 1024 + 8 * (2 * kWtStages + 1) + 16;
        rc = cuda_check(cudaFuncSetAttribute(k_dense_wgrad_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                        "cudaFuncSetAttribute(k_dense_wgrad_tma)");
        if (rc) return rc;
        const long long tiles = (n + kDtRows - 1) / kDtRows;
        grid = sm_count();
        if (tiles < grid) grid = tiles;
        k_dense_wgrad_tma<<<(unsigned)grid, kWtThreads, smem, st>>>(tm_g, tm_x, n_in, tiles, workspace, status);
        rc = after_launch("k_dense_wgrad_tma");
        if (rc) return rc;
    }
    k_dense_wgrad_tma_reduce<<<(kWtPartial + 255) / 256, 256, 0, st>>>(workspace, (int)grid, n_out, n_in, dw, db, accumulate);
    return after_launch("k_dense_wgrad_tma_reduce");
}
