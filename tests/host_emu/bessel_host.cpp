// TEST INFRASTRUCTURE: qbold_vi_b200/csrc/bessel.cuh -- the Bessel kernels of the hot path -- compiled for the HOST
// (g++ -DQB_HOST_EMU), so the CPU suite can check the very source the GPU kernels are built from against scipy.
// __device__ / __forceinline__ become plain inline, the bit casts become memcpy, and the header itself replaces its
// PTX statements (rsqrt.approx, mov.b64, fma/mul/add.rn.f32x2) by their IEEE meaning under QB_HOST_EMU.
#include <cmath>
#include <cstring>

#define __device__
#define __host__
#define __forceinline__ inline
static inline unsigned __float_as_uint(float f) {
    unsigned u;
    std::memcpy(&u, &f, 4);
    return u;
}
static inline float __uint_as_float(unsigned u) {
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}

#include "bessel.cuh"

// range: 0 = the production selection (bessel_pair), 1 / 2 / 3 = force the small / mid / big kernel (a warp pass whose
// arguments straddle a boundary runs ONE kernel on all lanes, so mid must hold on [2, 9] and big from 6.5).
extern "C" void qb_emu_bessel(const float* x, int n, int range, float* omj0, float* j1) {
    for (int i = 0; i < n; ++i) {
        switch (range) {
            case 1: qb::bessel_small<true>(x[i], omj0[i], j1[i]); break;
            case 2: qb::bessel_mid<true>(x[i], omj0[i], j1[i]); break;
            case 3: qb::bessel_big<true>(x[i], omj0[i], j1[i]); break;
            default: qb::bessel_pair<true>(x[i], omj0[i], j1[i]);
        }
    }
}

// The packed (FFMA2) accumulation steps on pairs (x[2i], x[2i+1]) with weight w and m = x / A:
// acc_i = w (1 - J0(x)),  acc_b = (w m) J1(x)  (small range: acc_b = w z S1(z) = w x J1(x), the kernel rescales by 1/A).
extern "C" void qb_emu_acc2(const float* x, int n_pairs, int range, float w, float A, float* acc_i, float* acc_b) {
    using namespace qb;
    for (int i = 0; i < n_pairs; ++i) {
        const f32x2 xx = pk2(x[2 * i], x[2 * i + 1]);
        const f32x2 mm = pk2(x[2 * i] / A, x[2 * i + 1] / A);
        f32x2 ai = pk1(0.f), ab = pk1(0.f);
        if (range == 1) acc_small2<true>(xx, pk1(w), ai, ab);
        else if (range == 2) acc_mid2<true>(xx, mm, pk1(w), ai, ab);
        else acc_big2<true>(xx, mm, pk1(w), ai, ab);
        upk2(ai, acc_i[2 * i], acc_i[2 * i + 1]);
        upk2(ab, acc_b[2 * i], acc_b[2 * i + 1]);
    }
}
