// TEST INFRASTRUCTURE: qbold_vi_b200/csrc/rng.cuh compiled for the host.  Its integer side (Philox4x32-10, the 24-bit
// uniform) is declared __host__ __device__ in the source, so the CPU suite can compare the very code the kernels run with
// oracle/philox.py word for word; the Box-Muller halves use device intrinsics and stay under __CUDACC__.
#include <stdint.h>

#define __device__
#define __host__
#define __forceinline__ inline

#include "rng.cuh"

extern "C" void qb_emu_philox(const uint64_t* index, int n, uint32_t stream, uint64_t seed, uint32_t* words, float* u) {
    for (int i = 0; i < n; ++i) {
        const qb::U4 r = qb::philox4x32_10((uint32_t)index[i], (uint32_t)(index[i] >> 32), stream, 0u, (uint32_t)seed,
                                           (uint32_t)(seed >> 32));
        words[4 * i + 0] = r.x;
        words[4 * i + 1] = r.y;
        words[4 * i + 2] = r.z;
        words[4 * i + 3] = r.w;
        u[4 * i + 0] = qb::u01(r.x);
        u[4 * i + 1] = qb::u01(r.y);
        u[4 * i + 2] = qb::u01(r.z);
        u[4 * i + 3] = qb::u01(r.w);
    }
}

// raw generator for the Random123 known-answer vectors (counter and key given word by word)
extern "C" void qb_emu_philox_raw(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
    const qb::U4 r = qb::philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
    out[0] = r.x;
    out[1] = r.y;
    out[2] = r.z;
    out[3] = r.w;
}

extern "C" uint32_t qb_emu_stream_id(int which) {
    switch (which) {
        case 0: return qb::kStreamReparam;
        case 1: return qb::kStreamKl;
        case 2: return qb::kStreamSnr;
        case 3: return qb::kStreamNoise;
        default: return qb::kStreamMisalign;
    }
}

// the keyed Feistel bijection of [0, n) that stands in for tf.random.shuffle (half_bits as qbold_generate derives it)
extern "C" void qb_emu_feistel(const uint64_t* i, int count, uint64_t n, uint64_t seed, uint64_t* out) {
    int bits = 1;
    while (bits < 64 && (1ull << bits) < n) ++bits;
    const int half_bits = (bits + 1) / 2;
    for (int k = 0; k < count; ++k) out[k] = qb::feistel_permute(i[k], n, half_bits, seed);
}
