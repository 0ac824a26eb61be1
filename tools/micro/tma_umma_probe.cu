// Probe: TMA (4-D tiled, SWIZZLE_128B) -> tcgen05.mma kind::tf32 with MN-major operands and a start address moved by
// whole 128-byte rows.  Prints, for a set of descriptor hypotheses, how many entries of D match the exact integer
// result.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/micro/tma_umma_probe tools/micro/tma_umma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool mbar_wait(unsigned bar, unsigned parity) {
    for (unsigned spin = 0; spin < (1u << 22); ++spin) {
        unsigned done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return true;
    }
    return false;
}
__device__ __forceinline__ uint64_t mk_desc(unsigned addr, unsigned lbo, unsigned sbo, unsigned base_off, unsigned layout) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)(base_off & 7u) << 49) | ((uint64_t)layout << 61);
}

struct Variant {
    unsigned lbo_a, sbo_a, lbo_b, sbo_b;   // bytes
    int row_off;                            // A start address moved by this many 128-byte rows
    int use_base_off;                       // put (row_off & 7) into the descriptor's base-offset field
    int a_mn, b_mn;                         // major bits
    int layout;                             // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B (TMA: 128B_ATOM_32B)
};

// x: [1, X, Y, 64], g: [1, X, Y, 64].  Loads 4 A boxes (lines xr-1 / xr, two channel halves, rows y0-1 .. y0+8) and 2 B
// boxes (line xr, rows y0 .. y0+7), dumps the raw shared memory, runs ONE M=128 N=64 K=8 MMA, dumps D.
__global__ void __launch_bounds__(128) k_probe(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_g,
                                               Variant v, int xr, int y0, float* smem_dump, float* d_out, int* status) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const unsigned raw = s32(smem_raw), base = (raw + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - raw);
    __shared__ uint64_t bars[2];
    __shared__ unsigned tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned LA = 2048, LB = 1024;                      // 16 rows reserved per A box, 8 per B box
    for (int i = tid; i < (4 * 2048 + 2 * 1024 + 4096) / 4; i += 128) reinterpret_cast<float*>(sm)[i] = -777.f;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_slot)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = tmem_slot;
    bool ok = true;
    if (tid == 0) {
        const unsigned tx = 4 * 10 * 128 + 2 * 8 * 128;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bars[0])), "r"(tx) : "memory");
        for (int dx = 0; dx < 2; ++dx)
            for (int ch = 0; ch < 2; ++ch)
                asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                             ::"r"(base + (2 * dx + ch) * LA), "l"(&tm_x), "r"(32 * ch), "r"(y0 - 1), "r"(xr + dx - 1), "r"(0), "r"(s32(&bars[0])) : "memory");
        for (int ch = 0; ch < 2; ++ch)
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                         ::"r"(base + 4 * LA + ch * LB), "l"(&tm_g), "r"(32 * ch), "r"(y0), "r"(xr), "r"(0), "r"(s32(&bars[0])) : "memory");
    }
    ok = mbar_wait(s32(&bars[0]), 0) && ok;
    __syncthreads();
    for (int i = tid; i < (4 * 2048 + 2 * 1024) / 4; i += 128) smem_dump[i] = reinterpret_cast<float*>(sm)[i];
    __syncthreads();
    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)v.a_mn << 15) | ((unsigned)v.b_mn << 16) |
                               ((unsigned)(64 >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
        const uint64_t da = mk_desc(base + v.row_off * 128, v.lbo_a, v.sbo_a, v.use_base_off ? (unsigned)v.row_off : 0u, (unsigned)v.layout);
        const uint64_t db = mk_desc(base + 4 * LA, v.lbo_b, v.sbo_b, 0u, (unsigned)v.layout);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bars[1])) : "memory");
    }
    ok = mbar_wait(s32(&bars[1]), 0) && ok;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int part = 0; part < 4; ++part) {
        float a[16];
        const unsigned taddr = tmem + part * 16 + ((unsigned)(warp * 32) << 16);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=f"(a[0]), "=f"(a[1]), "=f"(a[2]), "=f"(a[3]), "=f"(a[4]), "=f"(a[5]), "=f"(a[6]), "=f"(a[7]), "=f"(a[8]),
                       "=f"(a[9]), "=f"(a[10]), "=f"(a[11]), "=f"(a[12]), "=f"(a[13]), "=f"(a[14]), "=f"(a[15]) : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) d_out[(warp * 32 + lane) * 64 + part * 16 + j] = a[j];
    }
    if (!ok) atomicExch(status, 1);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int X = 4, Y = 16, C = 64;
    std::vector<float> hx(X * Y * C), hg(X * Y * C);
    auto xv = [&](int xr, int y, int c) -> float { return (xr < 0 || xr >= X || y < 0 || y >= Y) ? 0.f : (float)(((xr * 7 + y * 3 + c * 5) % 13) - 6); };
    auto gv = [&](int xr, int y, int c) -> float { return (xr < 0 || xr >= X || y < 0 || y >= Y) ? 0.f : (float)(((xr * 5 + y * 11 + c * 3) % 9) - 4); };
    for (int xr = 0; xr < X; ++xr) for (int y = 0; y < Y; ++y) for (int c = 0; c < C; ++c) {
        hx[(xr * Y + y) * C + c] = xv(xr, y, c);
        hg[(xr * Y + y) * C + c] = gv(xr, y, c);
    }
    float *dx, *dg, *dump, *dout; int* status;
    CK(cudaMalloc(&dx, hx.size() * 4)); CK(cudaMalloc(&dg, hg.size() * 4));
    CK(cudaMalloc(&dump, 16384 * 4)); CK(cudaMalloc(&dout, 128 * 64 * 4)); CK(cudaMalloc(&status, 4));
    CK(cudaMemcpy(dx, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dg, hg.data(), hg.size() * 4, cudaMemcpyHostToDevice));
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    auto make = [&](CUtensorMap* tm, float* p, int box_y, CUtensorMapSwizzle swz) {
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Y, (cuuint64_t)X, 1};
        cuuint64_t str[3] = {(cuuint64_t)C * 4, (cuuint64_t)Y * C * 4, (cuuint64_t)X * Y * C * 4};
        cuuint32_t box[4] = {32, (cuuint32_t)box_y, 1, 1}, es[4] = {1, 1, 1, 1};
        CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, p, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
    };
    CUtensorMap tmx, tmg, tmx32, tmg32; make(&tmx, dx, 10, CU_TENSOR_MAP_SWIZZLE_128B); make(&tmg, dg, 8, CU_TENSOR_MAP_SWIZZLE_128B);
    make(&tmx32, dx, 10, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B); make(&tmg32, dg, 8, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
    const int xr = 1, y0 = 0;   // y0 - 1 = -1: the first A row is the zero padding
    const unsigned LA = 2048, LB = 1024;
    std::vector<Variant> vs;
    vs.push_back({0, 1024, 0, 1024, 0, 0, 0, 0, 2});               // K-major both (harness check)
    vs.push_back({128, 512, LB, 512, 0, 0, 1, 1, 1});            // M atoms = the same 32 channels moved by 0..3 rows (LBO = one row)
    vs.push_back({128, 512, LB, 512, 1, 0, 1, 1, 1});
    for (int ro = 0; ro < 3; ++ro)
        for (int bo = 0; bo < 2; ++bo) {
            if (ro == 0 && bo) continue;
            vs.push_back({LA, 512, LB, 512, ro, bo, 1, 1, 1});
            vs.push_back({512, LA, 512, LB, ro, bo, 1, 1, 1});   // LBO / SBO swapped
            vs.push_back({LA, 1024, LB, 1024, ro, bo, 1, 1, 1});
        }
    std::vector<float> hd(128 * 64), hs(16384);
    bool printed_dump = false;
    for (const Variant& v : vs) {
        CK(cudaMemset(status, 0, 4)); CK(cudaMemset(dout, 0xFF, 128 * 64 * 4));
        k_probe<<<1, 128, 32768>>>(v.layout == 1 ? tmx32 : tmx, v.layout == 1 ? tmg32 : tmg, v, xr, y0, dump, dout, status);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("variant lbo_a=%u sbo_a=%u row_off=%d base_off=%d: launch failed: %s\n", v.lbo_a, v.sbo_a, v.row_off, v.use_base_off, cudaGetErrorString(e)); return 1; }
        int st; CK(cudaMemcpy(&st, status, 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hd.data(), dout, hd.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hs.data(), dump, hs.size() * 4, cudaMemcpyDeviceToHost));
        if (v.layout == 1 && !printed_dump) {
            printed_dump = true;
            // check the TMA image: element (row r, channel c) of box (dx, ch) expected at r*128 + ((c/4) ^ (r&7))*16 + (c%4)*4
            int bad = 0, tot = 0;
            for (int dxi = 0; dxi < 2; ++dxi) for (int ch = 0; ch < 2; ++ch) for (int r = 0; r < 10; ++r) for (int c = 0; c < 32; ++c) {
                const int off = (2 * dxi + ch) * LA + r * 128 + (((c >> 3) ^ (r & 3)) << 5) + (c & 7) * 4;
                const float want = xv(xr + dxi - 1, y0 - 1 + r, 32 * ch + c);
                ++tot; if (hs[off / 4] != want) { if (bad < 5) printf("  tma x mismatch dx=%d ch=%d r=%d c=%d got %g want %g\n", dxi, ch, r, c, hs[off / 4], want); ++bad; }
            }
            for (int ch = 0; ch < 2; ++ch) for (int r = 0; r < 8; ++r) for (int c = 0; c < 32; ++c) {
                const int off = 4 * LA + ch * LB + r * 128 + (((c >> 3) ^ (r & 3)) << 5) + (c & 7) * 4;
                const float want = gv(xr, y0 + r, 32 * ch + c);
                ++tot; if (hs[off / 4] != want) { if (bad < 10) printf("  tma g mismatch ch=%d r=%d c=%d got %g want %g\n", ch, r, c, hs[off / 4], want); ++bad; }
            }
            printf("TMA image: %d / %d elements where the swizzle formula expects them\n", tot - bad, tot);
        }
        // expected D from the dumped shared-memory image under the canonical SWIZZLE_128B layouts:
        //   K-major : element (row r, k) at (r / 8) * SBO + (r % 8) * 128 + (((k / 4) ^ (r % 8)) * 16) + (k % 4) * 4
        //   MN-major: element (mn, k)    at (mn / 32) * LBO + k * 128 + ((((mn % 32) / 4) ^ (k % 8)) * 16) + (mn % 4) * 4
        // (the swizzle XOR uses the absolute row, so a start moved by row_off rows reads row k + row_off)
        auto elem = [&](int base_bytes, int mn_major, unsigned lbo, unsigned sbo, int idx, int k, int row_off) -> double {
            long off;
            if (mn_major) { const int kr = k + row_off + (lbo == 128 ? idx / 32 : 0); off = (lbo == 128 ? 0 : (long)(idx / 32) * lbo) + (long)kr * 128 + ((((idx % 32) / 8) ^ (kr % 4)) * 32) + (idx % 8) * 4; }
            else { const int r = idx + row_off; off = (long)(r / 8) * sbo + (r % 8) * 128 + ((((k / 4) ^ (r % 8))) * 16) + (k % 4) * 4; }
            off += base_bytes;
            if (off < 0 || off / 4 >= (long)hs.size()) return 0.0;
            return hs[off / 4];
        };
        int match = 0, nonzero = 0;
        for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) {
            double s = 0;
            for (int k = 0; k < 8; ++k)
                s += elem(0, v.a_mn, v.lbo_a, v.sbo_a, m, k, v.row_off) * elem(4 * LA, v.b_mn, v.lbo_b, v.sbo_b, n, k, 0);
            if (hd[m * 64 + n] == (float)s) ++match;
            if (hd[m * 64 + n] != 0.f) ++nonzero;
        }
        printf("layout=%d a_mn=%d b_mn=%d ", v.layout, v.a_mn, v.b_mn);
        printf("lbo_a=%4u sbo_a=%4u lbo_b=%4u sbo_b=%4u row_off=%d base_off_field=%d : %5d / 8192 match, %5d nonzero, timeout=%d, D[0][0..3]=%g %g %g %g\n",
               v.lbo_a, v.sbo_a, v.lbo_b, v.sbo_b, v.row_off, v.use_base_off, match, nonzero, st, hd[0], hd[1], hd[2], hd[3]);
    }
    return 0;
}
