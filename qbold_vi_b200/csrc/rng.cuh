// Counter-based RNG shared by the kernels and (bit-for-bit on the integer side) by
// oracle/philox.py: Philox4x32-10 keyed by the 64-bit seed, counter = (index_lo, index_hi,
// stream, 0) where index is the GLOBAL voxel index, so results do not depend on how voxels
// are sharded over GPUs or chunked over launches.
//
// TensorFlow's own Philox streams (tf.random.*, seeded by tf.random.set_seed(1),
// reference train.py:458) cannot be reproduced offline; "identical seeds" is realised as this
// documented scheme plus explicit-draw entry points.
#pragma once
#include <stdint.h>

namespace qb {

// stream ids
constexpr uint32_t kStreamReparam = 0;        // ReparamTrickLayer draw of the likelihood term
constexpr uint32_t kStreamKl = 0x100;         // + (sample index >> 1): Monte-Carlo samples, see mc_normal_pair
constexpr uint32_t kStreamSnr = 0x10000;      // noise model: per-voxel SNR
constexpr uint32_t kStreamNoise = 0x10001;    // + pair index: noise normals
constexpr uint32_t kStreamMisalign = 0x20000; // misalignment: x selection, y first misaligned image, (z, w) normals

struct U4 {
    uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                     uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0;
        c1 = lo1;
        c2 = n2;
        c3 = lo0;
        k0 += W0;
        k1 += W1;
    }
    return U4{c0, c1, c2, c3};
}

// 24-bit uniform in (0, 1): (top 24 bits + 0.5) * 2^-24 -- never 0 or 1.
__host__ __device__ __forceinline__ float u01(uint32_t r) { return ((float)(r >> 8) + 0.5f) * 5.9604644775390625e-08f; }

// Keyed bijection of [0, n): 4-round balanced Feistel network on 2*half_bits bits with
// cycle walking.  Stands in for tf.random.shuffle (signals.py:279) without materialising a
// permutation array in HBM (1e9-voxel generation, BASELINE config 5).  oracle/philox.py
// implements the identical function.
__host__ __device__ __forceinline__ uint32_t feistel_round(uint32_t r, uint32_t key) {
    uint32_t h = r * 0x9E3779B1u + key;
    h ^= h >> 15;
    h *= 0x85EBCA77u;
    h ^= h >> 13;
    h *= 0xC2B2AE3Du;
    h ^= h >> 16;
    return h;
}

__host__ __device__ __forceinline__ uint64_t feistel_permute(uint64_t i, uint64_t n, int half_bits, uint64_t seed) {
    const uint32_t mask = (half_bits >= 32) ? 0xffffffffu : ((1u << half_bits) - 1u);
    uint64_t x = i;
    do {
        uint32_t l = (uint32_t)(x >> half_bits) & mask, r = (uint32_t)x & mask;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t key = (uint32_t)(seed >> (16 * (k & 1))) + 0x7F4A7C15u * (uint32_t)(k + 1) +
                                 (uint32_t)(seed >> 32);
            const uint32_t t = l ^ (feistel_round(r, key) & mask);
            l = r;
            r = t;
        }
        x = ((uint64_t)l << half_bits) | r;
    } while (x >= n);
    return x;
}

#if defined(__CUDACC__) || defined(QB_HOST_EMU)
// Two independent N(0,1) draws from two words (Box-Muller, accurate logf/sincospif).
__device__ __forceinline__ void box_muller(uint32_t r0, uint32_t r1, float& n0, float& n1) {
    const float rad = sqrtf(-2.0f * logf(u01(r0)));
    float s, c;
    sincospif(2.0f * u01(r1), &s, &c);
    n0 = rad * c;
    n1 = rad * s;
}

__device__ __forceinline__ void normal_pair(uint64_t seed, uint64_t index, uint32_t stream, float& n0, float& n1) {
    const U4 r = philox4x32_10((uint32_t)index, (uint32_t)(index >> 32), stream, 0u, (uint32_t)seed,
                               (uint32_t)(seed >> 32));
    box_muller(r.x, r.y, n0, n1);
}

// ---- Monte-Carlo sample draws (KL estimator, posterior statistics, likelihood map) ----------------------------
// Sample s of a voxel comes from Philox call (index, kStreamKl + (s >> 1)): words (x, y) for even s, (z, w) for odd
// s, so one call serves two samples.  These draws only feed sample averages, so Box-Muller runs on the SFU
// approximations (lg2 / sqrt / sin / cos .approx): absolute error of a draw <~ 2e-6 (up to 6e-4 for the 0.08 % of
// draws with radius < 0.04, where the 2^-22 absolute error of lg2.approx near 1 dominates) -- far below the
// estimators' sampling noise; four MUFU + ~8 FP32 instructions instead of ~75.  The likelihood's own reparameterisation draw and the dataset noise keep the
// accurate logf / sincospif path above.
#ifdef QB_HOST_EMU   // tests/host_emu: the SFU approximations replaced by the functions they approximate
__device__ __forceinline__ float lg2_approx(float x) { return log2f(x); }
__device__ __forceinline__ float sqrt_approx(float x) { return sqrtf(x); }
__device__ __forceinline__ float sin_approx(float x) { return sinf(x); }
__device__ __forceinline__ float cos_approx(float x) { return cosf(x); }
#else
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sin_approx(float x) {
    float y;
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float cos_approx(float x) {
    float y;
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
#endif   // QB_HOST_EMU

__device__ __forceinline__ void mc_box_muller(uint32_t r0, uint32_t r1, float& n0, float& n1) {
    const float rad = sqrt_approx(fmaxf(lg2_approx(u01(r0)) * -1.3862943611198906f, 0.f));   // sqrt(-2 ln u)
    const float x = fmaf(u01(r1), 6.283185307179586f, -3.141592653589793f);                   // angle - pi
    n0 = rad * -cos_approx(x);                                                                // cos(x + pi) = -cos x
    n1 = rad * -sin_approx(x);
}

__device__ __forceinline__ U4 mc_words(uint64_t seed, uint64_t index, int s) {
    return philox4x32_10((uint32_t)index, (uint32_t)(index >> 32), kStreamKl + (uint32_t)(s >> 1), 0u, (uint32_t)seed,
                         (uint32_t)(seed >> 32));
}

__device__ __forceinline__ void mc_normal_pair(const U4& r, int s, float& n0, float& n1) {
    const bool odd = (s & 1) != 0;
    mc_box_muller(odd ? r.z : r.x, odd ? r.w : r.y, n0, n1);
}

__device__ __forceinline__ void mc_normal_pair(uint64_t seed, uint64_t index, int s, float& n0, float& n1) {
    mc_normal_pair(mc_words(seed, index, s), s, n0, n1);
}
#endif

}  // namespace qb
