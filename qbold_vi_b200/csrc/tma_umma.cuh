// TMA + tcgen05 building blocks shared by the encoder's tensor-core kernels (csrc/encoder_conv_tma.cu,
// csrc/encoder_dense_tma.cu): mbarriers with bounded waits, tiled TMA loads / stores, shared-memory matrix descriptors,
// the kind::tf32 instruction descriptor, TMEM loads, and the host-side tensor-map encoder.
#pragma once
#include <cuda.h>

#include "launch.h"

namespace qb {

__device__ __forceinline__ unsigned ct_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ct_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void ct_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void ct_mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void ct_mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait (a descriptor or byte-count mistake must not wedge the GPU); false on timeout.  The polling loop lives
// inside ONE asm statement: a C-level loop that branches on the (per-thread) try_wait result makes everything after it
// look divergent to the compiler, which then refuses to keep the MMA descriptors in uniform registers.
__device__ __forceinline__ bool ct_mbar_wait(unsigned bar, unsigned parity) {
    unsigned done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .u32 n;\n\t"
        "mov.u32 n, 0;\n\t"
        "CT_WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "@p bra CT_WAIT_DONE;\n\t"
        "add.u32 n, n, 1;\n\t"
        "setp.lt.u32 p, n, 0x1000000;\n\t"
        "@p bra CT_WAIT_LOOP;\n\t"
        "mov.u32 %0, 0;\n\t"
        "bra CT_WAIT_EXIT;\n\t"
        "CT_WAIT_DONE:\n\t"
        "mov.u32 %0, 1;\n\t"
        "CT_WAIT_EXIT:\n\t"
        "}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
}
// The same for threads that are NOT on the critical path (epilogue warps parked until the accumulators are final, a
// producer several stages ahead): sleep between polls so the polling does not compete with the tensor pipe's operand
// fetches for shared-memory bandwidth.
__device__ __forceinline__ bool ct_mbar_wait_relaxed(unsigned bar, unsigned parity, unsigned sleep_ns) {
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
        __nanosleep(sleep_ns);
    }
    return false;
}
// One lane of a converged warp.  The MMA / TMA issue loops are run by ALL lanes of their warp with warp-uniform values and
// only the instruction itself sits under this predicate: tcgen05.mma and the TMA instructions take their operands from
// uniform registers, and a loop that runs on a single lane (inside `if (lane == 0)`) makes the compiler move every
// descriptor through an ELECT / R2UR.BROADCAST "waterfall" -- ~90 clk per MMA instead of the 48 the tensor pipe needs.
__device__ __forceinline__ bool ct_elect_one() {
    unsigned pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void ct_commit(unsigned bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void ct_mma(unsigned tmem_d, uint64_t desc_a, uint64_t desc_b, unsigned idesc, unsigned acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc)
        : "memory");
}
// 4-D tiled TMA load, completion on an mbarrier (bytes of the whole box, zero-filled parts included).
__device__ __forceinline__ void ct_tma_4d(unsigned dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3,
                                          unsigned bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], "
        "[%6];"
        ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
        : "memory");
}
// Shared-memory matrix descriptor: start address, leading / stride byte offsets (16-byte units), descriptor version 1
// (Blackwell), layout type (2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B: 32-byte chunks XOR (row & 3), the only
// swizzled layout tcgen05 accepts for MN-major tf32 operands; TMA writes it as CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B).
// Verified on a B200 with tools/micro/tma_umma_probe.cu, including start addresses moved by whole 128-byte rows.
constexpr unsigned kLayoutSw128 = 2, kLayoutSw128Base32 = 1;
__device__ __forceinline__ uint64_t ct_desc(unsigned addr, unsigned lbo_bytes, unsigned sbo_bytes, unsigned layout) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46) | ((uint64_t)layout << 61);
}
// D = F32, A = B = TF32; majors: bit 15 (A) / 16 (B) set = MN-major; N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ unsigned ct_idesc(int m, int n, bool a_mn, bool b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((unsigned)(n >> 3) << 17) | ((unsigned)(m >> 4) << 24);
}
__device__ __forceinline__ void ct_tmem_ld16(unsigned taddr, float* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
        "%15}, [%16];"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
          "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 2-D tiled TMA load / store (store: bulk async-group completion)
__device__ __forceinline__ void ct_tma_2d(unsigned dst, const CUtensorMap* tm, int c0, int c1, unsigned bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void ct_tma_store_2d(const CUtensorMap* tm, int c0, int c1, unsigned src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(tm), "r"(c0), "r"(c1), "r"(src)
                 : "memory");
}
__device__ __forceinline__ void ct_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void ct_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ct_mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void ct_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn();                 // cuTensorMapEncodeTiled through cudaGetDriverEntryPoint, or nullptr
// Tensor map over a z-outer activation [n_images, X, Y, C] float32: dims (C, Y, X, n_images), box 32 channels x box_y
// voxels of one image line, zero fill outside.
int activation_map(CUtensorMap* tm, const float* ptr, int c, long long n_images, int X, int Y, int box_y,
                   CUtensorMapSwizzle swizzle);
// Tensor map over a row-major matrix [rows, cols] float32 (cols * 4 a multiple of 16): box 32 columns x box_rows rows.
int matrix_map(CUtensorMap* tm, const float* ptr, long long rows, int cols, int box_rows, CUtensorMapSwizzle swizzle);

}  // namespace qb
