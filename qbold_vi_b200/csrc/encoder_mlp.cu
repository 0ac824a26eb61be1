// Stream 1 of the amortization network (reference model.py:122-223: normalise_data -> 1x1x1 conv -> ReLU ->
// [1x1x1 conv -> ReLU] x n_blocks -> 1x1x1 conv), i.e. a per-voxel MLP n_in -> H -> ... -> H -> n_out, as ONE
// sm_100a kernel on the 5th-generation tensor cores.  Forward / inference only: it produces the
// voxel-wise posterior the fine-tuning stage uses as its prior (train.py:26-31) and save_predictions maps.
//
// Mapping: a group of 128 threads owns tiles of 128 voxels = the 128 TMEM lanes; a CTA runs kGroups independent
// groups (own A tile, TMEM columns, mbarrier and named barrier; the weights are shared) so that one group's
// epilogue overlaps the others' tensor-core work -- the per-layer chain ld -> epilogue -> fence -> mma -> commit
// is latency-, not throughput-bound.  Activations (A, 128 x 64 fp32) and all layer
// weights (B, [out, in] = K-major, padded to 64) live in shared memory in the canonical K-major SWIZZLE_128B
// layout; thread 0 issues tcgen05.mma kind::tf32 (fp32 bit patterns in, fp32 accumulate in TMEM), completion
// comes back through tcgen05.commit -> mbarrier, each thread then pulls its own voxel's 64 accumulators with
// tcgen05.ld 32x32b, applies bias + ReLU and writes the next layer's A row.  Nothing but the input images and the
// n_out outputs touches HBM (44 B + 20 B per voxel for optimal.yaml).
#include "launch.h"

namespace qb {

namespace {

constexpr int kTile = 128;          // voxels per tile = TMEM lanes = threads per CTA
constexpr int kH = 64;              // padded hidden width (N of the hidden layers, K of the next)
constexpr int kNOut = 16;           // padded output width
constexpr int kMaxMid = 6;          // hidden->hidden layers
constexpr int kKBlock = 32;         // floats per 128-byte swizzle row
constexpr int kGroups = 4;          // independent 128-thread tile pipelines per CTA (1 CTA / SM: 128 + 46 KB smem)
constexpr int kTmemCols = 64 * kGroups;

constexpr int kATileFloats = 2 * kTile * kKBlock;       // 2 K-blocks x 128 rows x 32 floats = 32 KB
constexpr int kW0Floats = kH * kKBlock;                 // 1 K-block  x 64 rows            = 8 KB
constexpr int kWmFloats = 2 * kH * kKBlock;             // 2 K-blocks x 64 rows            = 16 KB
constexpr int kWoFloats = 2 * kNOut * kKBlock;          // 2 K-blocks x 16 rows            = 4 KB

// float index of element (row, k) inside a K-major SWIZZLE_128B tile with `rows` rows: 8-row x 128-byte atoms,
// the 16-byte chunk index XOR-ed with the row index inside the atom; K-blocks of 32 floats are separate slabs.
__host__ __device__ __forceinline__ int swz(int rows, int row, int k) {
    const int kb = k >> 5, kk = k & 31;
    return kb * (rows * kKBlock) + (row >> 3) * 256 + (row & 7) * 32 + ((((kk >> 2) ^ (row & 7))) << 2) + (kk & 3);
}

struct MlpWeights {
    const float* w_in;              // [H, n_in]
    const float* b_in;              // [H]
    const float* w_mid[kMaxMid];    // [H, H]
    const float* b_mid[kMaxMid];
    const float* w_out;             // [n_out, H]
    const float* b_out;
    int n_in, n_hidden, n_mid, n_out;
};

__host__ __device__ __forceinline__ int blob_floats(int n_mid) {
    return kW0Floats + n_mid * kWmFloats + kWoFloats + (n_mid + 1) * kH + kNOut;
}

// Packs torch-layout weights into the shared-memory image the MLP kernel copies verbatim (zero padded).
__global__ void k_mlp_pack(MlpWeights w, float* __restrict__ blob) {
    const int total = blob_floats(w.n_mid);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) blob[i] = 0.f;
}

__global__ void k_mlp_fill(MlpWeights w, float* __restrict__ blob) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (int i = tid; i < w.n_hidden * w.n_in; i += nth) {
        const int r = i / w.n_in, k = i % w.n_in;
        blob[swz(kH, r, k)] = w.w_in[i];
    }
    float* mid = blob + kW0Floats;
    for (int l = 0; l < w.n_mid; ++l)
        for (int i = tid; i < w.n_hidden * w.n_hidden; i += nth) {
            const int r = i / w.n_hidden, k = i % w.n_hidden;
            mid[l * kWmFloats + swz(kH, r, k)] = w.w_mid[l][i];
        }
    float* wo = mid + w.n_mid * kWmFloats;
    for (int i = tid; i < w.n_out * w.n_hidden; i += nth) {
        const int r = i / w.n_hidden, k = i % w.n_hidden;
        wo[swz(kNOut, r, k)] = w.w_out[i];
    }
    float* bias = wo + kWoFloats;
    for (int i = tid; i < w.n_hidden; i += nth) {
        bias[i] = w.b_in[i];
        for (int l = 0; l < w.n_mid; ++l) bias[(l + 1) * kH + i] = w.b_mid[l][i];
    }
    for (int i = tid; i < w.n_out; i += nth) bias[(w.n_mid + 1) * kH + i] = w.b_out[i];
}

// ---- PTX wrappers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void group_sync(int group) {        // named barrier 1 + group, 128 threads
    asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(kTile) : "memory");
}

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

// Bounded wait: a descriptor mistake must not wedge the GPU; on timeout the caller records an error flag.
__device__ __forceinline__ bool mbar_wait(unsigned bar, unsigned parity) {
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

__device__ __forceinline__ void umma_commit(unsigned bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc]^T, kind::tf32, issued by one thread
__device__ __forceinline__ void umma_tf32(unsigned tmem_d, uint64_t desc_a, uint64_t desc_b, unsigned idesc,
                                          unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor: start address (16-byte units), stride between 8-row
// atoms = 1024 B, descriptor version 1 (sm_100), layout type 2.
__device__ __forceinline__ uint64_t smem_desc(unsigned addr) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(1024u >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}

// Instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128, N = n.
__device__ __forceinline__ unsigned instr_desc(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(n >> 3) << 17) | ((unsigned)(kTile >> 4) << 24);
}

__device__ __forceinline__ void tmem_ld16(unsigned taddr, float* v) {
    unsigned r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
        "%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// The destination registers are only defined after tcgen05.wait::ld: they are bound straight to the caller's
// floats so that no instruction touches them in between.
__device__ __forceinline__ void tmem_ld16_nowait(unsigned taddr, float* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
        "%15}, [%16];"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
          "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void sts_f4(unsigned addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// byte address of 16-byte chunk `chunk` (0..15 over the two K-blocks) of row `row` in the A tile
__device__ __forceinline__ unsigned a_chunk_addr(unsigned a_base, int row, int chunk) {
    const int kb = chunk >> 3, c = chunk & 7;
    return a_base + kb * (kTile * 128) + (row >> 3) * 1024 + (row & 7) * 128 + ((c ^ (row & 7)) << 4);
}

}  // namespace

// data [n, n_in] raw images -> q [n, n_out].  status[0] is set non-zero if a tensor-core completion never arrived.
__global__ void __launch_bounds__(kTile * kGroups, 1) k_encoder_mlp(const float* __restrict__ data, const float* __restrict__ blob,
                                                       int n_in, int n_mid, int n_out, int se_idx, int multi_norm,
                                                       int64_t n, float* __restrict__ q, int* __restrict__ status) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // 1024-byte alignment of the swizzle atoms (dynamic shared memory starts 1024-aligned when nothing static
    // precedes it; align explicitly anyway)
    const unsigned raw = smem_u32(smem_raw);
    const unsigned base = (raw + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - raw);
    float* sA = reinterpret_cast<float*>(sm);
    float* sW = sA + kGroups * kATileFloats;
    const int wfloats = blob_floats(n_mid);
    float* sBias = sW + kW0Floats + n_mid * kWmFloats + kWoFloats;
    uint64_t* sBar = reinterpret_cast<uint64_t*>(sW + ((wfloats + 3) & ~3));
    unsigned* sTmem = reinterpret_cast<unsigned*>(sBar + kGroups);

    const int warp = threadIdx.x >> 5, group = threadIdx.x / kTile, tid = threadIdx.x % kTile;
    const unsigned a_base = base + group * (kATileFloats * 4), w_base = base + kGroups * kATileFloats * 4;
    const unsigned bar = smem_u32(sBar + group);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sTmem)),
                     "r"((unsigned)kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {   // weights + biases: verbatim copy of the packed image
        const float4* src = reinterpret_cast<const float4*>(blob);
        float4* dst = reinterpret_cast<float4*>(sW);
        for (int i = threadIdx.x; i < wfloats / 4; i += kTile * kGroups) dst[i] = __ldg(src + i);
    }
    fence_async_smem();
    tc_before_sync();
    __syncthreads();
    tc_after_sync();
    const unsigned tmem_all = *sTmem;
    const unsigned tmem = tmem_all + group * 64;                          // this group's 64 accumulator columns
    const unsigned taddr = tmem + ((unsigned)((warp & 3) * 32) << 16);    // a warp reaches lanes 32 (warp % 4) ...
    const unsigned idesc_h = instr_desc(kH), idesc_o = instr_desc(kNOut);
    const int ks_in = (n_in + 7) >> 3;                                    // k-steps of 8 floats for the first layer
    unsigned phase = 0;
    bool ok = true;

    const int64_t tiles = (n + kTile - 1) / kTile;
    for (int64_t tile = (int64_t)blockIdx.x * kGroups + group; tile < tiles; tile += (int64_t)gridDim.x * kGroups) {
        const int64_t v = tile * kTile + tid;
        // ---- normalise_data (model.py:97-113): clip, divide by the tau = 0 image (or the 3-image mean), log
        {
            const float* row = data + (v < n ? v : n - 1) * n_in;
            float ref = 0.f;
            if (multi_norm) {
                for (int i = se_idx - 1; i <= se_idx + 1; ++i) ref += fminf(fmaxf(__ldg(row + i), 1e-2f), 1e8f);
                ref = ref / 3.0f;
            } else {
                ref = fminf(fmaxf(__ldg(row + se_idx), 1e-2f), 1e8f);
            }
            // log(x / ref) as log x - log ref: no per-element division (difference < 1e-6 absolute, the operands are
            // rounded to TF32 right after)
            const float log_ref = logf(ref);
            for (int ks = 0; ks < ks_in; ++ks) {                       // 8 inputs = one k-step = two 16-byte chunks
                float x[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int k = ks * 8 + i;
                    x[i] = 0.f;
                    if (k < n_in) x[i] = logf(fminf(fmaxf(__ldg(row + k), 1e-2f), 1e8f)) - log_ref;
                }
                sts_f4(a_chunk_addr(a_base, tid, 2 * ks), x[0], x[1], x[2], x[3]);
                sts_f4(a_chunk_addr(a_base, tid, 2 * ks + 1), x[4], x[5], x[6], x[7]);
            }
        }
        tc_before_sync();
        fence_async_smem();
        group_sync(group);
        if (tid == 0) {
            tc_after_sync();
            for (int k = 0; k < ks_in; ++k)
                umma_tf32(tmem, smem_desc(a_base) + 2 * k, smem_desc(w_base) + 2 * k, idesc_h, k > 0);
            umma_commit(bar);
        }
        ok = mbar_wait(bar, phase) && ok;
        phase ^= 1;
        tc_after_sync();

        // ---- hidden layers: bias + ReLU epilogue writes the next A tile, then 8 k-steps over the two K-blocks
        for (int l = 0; l <= n_mid; ++l) {
            const float* bias = sBias + l * kH;
            float acc[64];
#pragma unroll
            for (int part = 0; part < 4; ++part) tmem_ld16_nowait(taddr + part * 16, acc + part * 16);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const float4 b = *reinterpret_cast<const float4*>(bias + c * 4);
                sts_f4(a_chunk_addr(a_base, tid, c), fmaxf(acc[c * 4 + 0] + b.x, 0.f), fmaxf(acc[c * 4 + 1] + b.y, 0.f),
                       fmaxf(acc[c * 4 + 2] + b.z, 0.f), fmaxf(acc[c * 4 + 3] + b.w, 0.f));
            }
            tc_before_sync();
            fence_async_smem();
            group_sync(group);
            const bool last = (l == n_mid);
            if (tid == 0) {
                tc_after_sync();
                const unsigned wl = w_base + (kW0Floats + l * kWmFloats) * 4;
                const unsigned kb_stride_w = (last ? kNOut : kH) * 128;
                for (int k = 0; k < 8; ++k) {
                    const unsigned ao = a_base + (k >> 2) * (kTile * 128), wo = wl + (k >> 2) * kb_stride_w;
                    umma_tf32(tmem, smem_desc(ao) + 2 * (k & 3), smem_desc(wo) + 2 * (k & 3), last ? idesc_o : idesc_h,
                              k > 0);
                }
                umma_commit(bar);
            }
            ok = mbar_wait(bar, phase) && ok;
            phase ^= 1;
            tc_after_sync();
        }
        // ---- output layer: n_out <= 16 accumulators + bias -> HBM
        {
            float acc[16];
            tmem_ld16(taddr, acc);
            const float* bias = sBias + (n_mid + 1) * kH;
            if (v < n) {
#pragma unroll
                for (int j = 0; j < kNOut; ++j)
                    if (j < n_out) q[v * n_out + j] = acc[j] + bias[j];
            }
        }
    }
    if (!ok && status != nullptr) atomicExch(status, 1);
    tc_before_sync();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_all), "r"((unsigned)kTmemCols)
                     : "memory");
    }
}

}  // namespace qb

using namespace qb;

extern "C" int qbold_encoder_mlp_blob_floats(int32_t n_mid) {
    return (n_mid < 1 || n_mid > kMaxMid) ? -1 : blob_floats(n_mid);
}

extern "C" int qbold_encoder_mlp_pack(const float* w_in, const float* b_in, const float* const* w_mid,
                                      const float* const* b_mid, const float* w_out, const float* b_out, int32_t n_in,
                                      int32_t n_hidden, int32_t n_mid, int32_t n_out, float* blob, void* stream) {
    if (n_in < 1 || n_in > 32 || n_hidden < 1 || n_hidden > kH || n_mid < 1 || n_mid > kMaxMid || n_out < 1 ||
        n_out > kNOut)
        return fail(QBOLD_EUNSUPPORTED,
                    "qbold_encoder_mlp_pack: supports n_in <= 32, hidden <= 64, 1..%d hidden->hidden layers, n_out <= 16",
                    kMaxMid);
    if (!w_in || !b_in || !w_mid || !b_mid || !w_out || !b_out || !blob)
        return fail(QBOLD_EINVAL, "qbold_encoder_mlp_pack: null pointer");
    MlpWeights w{};
    w.w_in = w_in;
    w.b_in = b_in;
    for (int l = 0; l < n_mid; ++l) {
        if (!w_mid[l] || !b_mid[l]) return fail(QBOLD_EINVAL, "qbold_encoder_mlp_pack: null layer pointer");
        w.w_mid[l] = w_mid[l];
        w.b_mid[l] = b_mid[l];
    }
    w.w_out = w_out;
    w.b_out = b_out;
    w.n_in = n_in;
    w.n_hidden = n_hidden;
    w.n_mid = n_mid;
    w.n_out = n_out;
    k_mlp_pack<<<32, 256, 0, (cudaStream_t)stream>>>(w, blob);
    int rc = after_launch("k_mlp_pack");
    if (rc) return rc;
    k_mlp_fill<<<32, 256, 0, (cudaStream_t)stream>>>(w, blob);
    return after_launch("k_mlp_fill");
}

extern "C" int qbold_encoder_mlp_forward(const float* data, const float* blob, int32_t n_in, int32_t n_mid,
                                         int32_t n_out, int32_t se_idx, int32_t multi_image_normalisation, int64_t n,
                                         float* q, int32_t* status, void* stream) {
    if (n_in < 1 || n_in > 32 || n_mid < 1 || n_mid > kMaxMid || n_out < 1 || n_out > kNOut || n < 0 || se_idx < 0 ||
        se_idx >= n_in || (multi_image_normalisation && (se_idx < 1 || se_idx + 1 >= n_in)))
        return fail(QBOLD_EINVAL, "qbold_encoder_mlp_forward: bad argument");
    if (n == 0) return QBOLD_OK;
    if (!data || !blob || !q) return fail(QBOLD_EINVAL, "qbold_encoder_mlp_forward: null pointer");
    const size_t smem = 1024 + (size_t)(kGroups * kATileFloats + ((blob_floats(n_mid) + 3) & ~3)) * 4 + 8 * kGroups + 16;
    int rc = cuda_check(cudaFuncSetAttribute(k_encoder_mlp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                        "cudaFuncSetAttribute(k_encoder_mlp)");      // per device; cheap, so not cached
    if (rc) return rc;
    const int64_t want = ((n + kTile - 1) / kTile + kGroups - 1) / kGroups;
    const int64_t cap = (int64_t)sm_count();
    k_encoder_mlp<<<(unsigned)(want < cap ? want : cap), kTile * kGroups, smem, (cudaStream_t)stream>>>(
        data, blob, n_in, n_mid, n_out, se_idx, multi_image_normalisation, n, q, status);
    return after_launch("k_encoder_mlp");
}

// ---------------------------------------------------------------------------------------------------------------------
// One Dense layer on the same machinery, for the TRAINING passes of the encoder:
//     Y[n, n_out] = act( (X[n, n_in] (* [M > 0])) * B^T + bias ),   B [n_out, n_in] row-major (= K-major)
// forward: B = W, bias, optional ReLU;  input gradient: X = dY, M = the layer's ReLU output (fuses ReLU'),
// B = W^T (packed by k_dense_pack with transpose = 1), no bias.  n_in, n_out <= 64 and multiples of 4.
// HBM-bound (240 B in + 240 B out per voxel for 60 channels); four tile pipelines per CTA as above.
namespace qb {

namespace {
constexpr int kDenseWFloats = 2 * kH * kKBlock;      // one 64 x 64 weight tile, 16 KB
}

// tile[swz(64, r, k)] = transpose ? w[k * ld + r] : w[r * ld + k]   (r < rows, k < cols; rest zero); bias appended
__global__ void k_dense_pack(const float* __restrict__ w, const float* __restrict__ bias, int rows, int cols,
                             int transpose, float* __restrict__ tile) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (int e = tid; e < kH * kH; e += nth) {
        const int r = e >> 6, k = e & 63;
        float v = 0.f;
        if (r < rows && k < cols) v = transpose ? w[k * rows + r] : w[r * cols + k];
        tile[swz(kH, r, k)] = v;
    }
    for (int e = tid; e < kH; e += nth) tile[kDenseWFloats + e] = (bias != nullptr && e < rows) ? bias[e] : 0.f;
}

__global__ void __launch_bounds__(kTile * kGroups, 1) k_dense_tc(const float* __restrict__ x, const float* __restrict__ relu_mask,
                                                                 const float* __restrict__ packed, int n_in, int n_out,
                                                                 int relu, int64_t n, float* __restrict__ y,
                                                                 int* __restrict__ status) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const unsigned raw = smem_u32(smem_raw);
    const unsigned base = (raw + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - raw);
    float* sA = reinterpret_cast<float*>(sm);
    float* sW = sA + kGroups * kATileFloats;
    float* sBias = sW + kDenseWFloats;
    uint64_t* sBar = reinterpret_cast<uint64_t*>(sBias + kH);
    unsigned* sTmem = reinterpret_cast<unsigned*>(sBar + kGroups);

    const int warp = threadIdx.x >> 5, group = threadIdx.x / kTile, tid = threadIdx.x % kTile;
    const unsigned a_base = base + group * (kATileFloats * 4), w_base = base + kGroups * kATileFloats * 4;
    const unsigned bar = smem_u32(sBar + group);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sTmem)),
                     "r"((unsigned)kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        const float4* src = reinterpret_cast<const float4*>(packed);
        float4* dst = reinterpret_cast<float4*>(sW);
        for (int i = threadIdx.x; i < (kDenseWFloats + kH) / 4; i += kTile * kGroups) dst[i] = __ldg(src + i);
        float4* za = reinterpret_cast<float4*>(sA);                       // K padding of the A tiles stays zero
        for (int i = threadIdx.x; i < kGroups * kATileFloats / 4; i += kTile * kGroups)
            za[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    fence_async_smem();
    tc_before_sync();
    __syncthreads();
    tc_after_sync();
    const unsigned tmem_all = *sTmem;
    const unsigned tmem = tmem_all + group * 64;
    const unsigned taddr = tmem + ((unsigned)((warp & 3) * 32) << 16);
    const unsigned idesc = instr_desc(kH);
    const int chunks = n_in >> 2, ksteps = (n_in + 7) >> 3, ochunks = n_out >> 2;
    unsigned phase = 0;
    bool ok = true;

    const int64_t tiles = (n + kTile - 1) / kTile;
    for (int64_t tile = (int64_t)blockIdx.x * kGroups + group; tile < tiles; tile += (int64_t)gridDim.x * kGroups) {
        const int64_t v0 = tile * kTile;
        const int rows = (int)((n - v0 < kTile) ? (n - v0) : kTile);
        // ---- stage the A tile: coalesced 16-byte pieces of the contiguous [rows x n_in] span -> swizzled rows
        for (int e = tid; e < kTile * chunks; e += kTile) {
            const int r = e / chunks, c = e - r * chunks;
            const unsigned dst = a_chunk_addr(a_base, r, c);
            if (r < rows) {
                const float* src = x + (v0 + r) * n_in + c * 4;
                if (relu_mask == nullptr) {
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
                } else {
                    const float4 a = __ldg(reinterpret_cast<const float4*>(src));
                    const float4 m = __ldg(reinterpret_cast<const float4*>(relu_mask + (v0 + r) * n_in + c * 4));
                    sts_f4(dst, m.x > 0.f ? a.x : 0.f, m.y > 0.f ? a.y : 0.f, m.z > 0.f ? a.z : 0.f,
                           m.w > 0.f ? a.w : 0.f);
                }
            } else {
                sts_f4(dst, 0.f, 0.f, 0.f, 0.f);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        tc_before_sync();
        fence_async_smem();
        group_sync(group);
        if (tid == 0) {
            tc_after_sync();
            for (int k = 0; k < ksteps; ++k) {
                const unsigned ao = a_base + (k >> 2) * (kTile * 128), wo = w_base + (k >> 2) * (kH * 128);
                umma_tf32(tmem, smem_desc(ao) + 2 * (k & 3), smem_desc(wo) + 2 * (k & 3), idesc, k > 0);
            }
            umma_commit(bar);
        }
        ok = mbar_wait(bar, phase) && ok;
        phase ^= 1;
        tc_after_sync();
        // ---- epilogue: this thread's row, bias (+ ReLU), 16-byte stores
        float acc[64];
#pragma unroll
        for (int part = 0; part < 4; ++part) tmem_ld16_nowait(taddr + part * 16, acc + part * 16);
        tmem_ld_wait();
        if (tid < rows) {
            float* dst = y + (v0 + tid) * n_out;
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                if (c < ochunks) {
                    const float4 b = *reinterpret_cast<const float4*>(sBias + c * 4);
                    float4 o = make_float4(acc[c * 4] + b.x, acc[c * 4 + 1] + b.y, acc[c * 4 + 2] + b.z,
                                           acc[c * 4 + 3] + b.w);
                    if (relu) o = make_float4(fmaxf(o.x, 0.f), fmaxf(o.y, 0.f), fmaxf(o.z, 0.f), fmaxf(o.w, 0.f));
                    *reinterpret_cast<float4*>(dst + c * 4) = o;
                }
            }
        }
        tc_before_sync();            // the next tile's MMA overwrites these TMEM columns: order the loads before it
        group_sync(group);
    }
    if (!ok && status != nullptr) atomicExch(status, 1);
    tc_before_sync();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_all), "r"((unsigned)kTmemCols)
                     : "memory");
    }
}

}  // namespace qb

extern "C" int qbold_dense_tc_packed_floats(void) { return kDenseWFloats + kH; }

extern "C" int qbold_dense_tc_pack(const float* w, const float* bias, int32_t rows, int32_t cols, int32_t transpose,
                                   float* packed, void* stream) {
    // rows / cols describe the B operand [n_out, n_in] AFTER the optional transpose of the stored matrix
    if (rows < 1 || rows > kH || cols < 1 || cols > kH) return fail(QBOLD_EUNSUPPORTED, "qbold_dense_tc_pack: <= 64 x 64");
    if (!w || !packed) return fail(QBOLD_EINVAL, "qbold_dense_tc_pack: null pointer");
    k_dense_pack<<<16, 256, 0, (cudaStream_t)stream>>>(w, bias, rows, cols, transpose, packed);
    return after_launch("k_dense_pack");
}

extern "C" int qbold_dense_tc(const float* x, const float* relu_mask, const float* packed, int32_t n_in, int32_t n_out,
                              int32_t relu, int64_t n, float* y, int32_t* status, void* stream) {
    if (n_in < 4 || n_in > kH || (n_in & 3) || n_out < 4 || n_out > kH || (n_out & 3) || n < 0)
        return fail(QBOLD_EUNSUPPORTED, "qbold_dense_tc: n_in, n_out must be multiples of 4 in [4, 64] (got %d, %d)", n_in,
                    n_out);
    if (n == 0) return QBOLD_OK;
    if (!x || !packed || !y) return fail(QBOLD_EINVAL, "qbold_dense_tc: null pointer");
    if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 15) ||
        (relu_mask && (reinterpret_cast<uintptr_t>(relu_mask) & 15)))
        return fail(QBOLD_EINVAL, "qbold_dense_tc: x, y, relu_mask must be 16-byte aligned");
    const size_t smem = 1024 + (size_t)(kGroups * kATileFloats + kDenseWFloats + kH) * 4 + 8 * kGroups + 16;
    int rc = cuda_check(cudaFuncSetAttribute(k_dense_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                        "cudaFuncSetAttribute(k_dense_tc)");
    if (rc) return rc;
    const int64_t want = ((n + kTile - 1) / kTile + kGroups - 1) / kGroups;
    const int64_t cap = (int64_t)sm_count();
    k_dense_tc<<<(unsigned)(want < cap ? want : cap), kTile * kGroups, smem, (cudaStream_t)stream>>>(
        x, relu_mask, packed, n_in, n_out, relu, n, y, status);
    return after_launch("k_dense_tc");
}
