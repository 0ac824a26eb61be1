// Device-side FP32 evaluation of the pair (1 - J0(x), J1(x)) for x >= 0.
//
// The qBOLD tissue integrand (reference signals.py:169-171) needs 1 - J0, never J0
// itself, and TensorFlow's gradient of bessel_j0 is -bessel_j1.  Both are evaluated
// here without fast-math intrinsics (no __cosf/__expf/__fdividef):
//
//   x <= QB_X_SPLIT : 1 - J0 = z*P0(z), J1 = x*P1(z), z = x*x   (no cancellation for small x,
//                     where the quadrature weights ~ 1/u^2 are largest)
//   x >  QB_X_SPLIT : modulus/phase form  J_n = rsqrt(x) A_n(w) cos(x - (2n+1)pi/4 + q F_n(w)),
//                     q = 1/x = rsqrt(x)^2, w = q^2; the cosine is a polynomial in r^2 after a
//                     two-constant Cody-Waite reduction mod pi (|x| < ~1e4 keeps n*PI_HI exact).
//
// Coefficients: our own near-minimax fits (tools/fit_bessel.py -> bessel_coef.h); max abs
// error vs float64: 1-J0 2.0e-7 (1.6e-7 relative), J1 2.1e-7 on [0,3]; J0 2.6e-7, J1 3.7e-7
// on [3,32] (the floor there is the float32 rounding of the phase, as in the reference).
#pragma once
#include "bessel_coef.h"

namespace qb {

#define QB_H5(P, t) fmaf(fmaf(fmaf(fmaf(fmaf(QB_##P##_5, t, QB_##P##_4), t, QB_##P##_3), t, QB_##P##_2), t, QB_##P##_1), t, QB_##P##_0)
#define QB_H4(P, t) fmaf(fmaf(fmaf(fmaf(QB_##P##_4, t, QB_##P##_3), t, QB_##P##_2), t, QB_##P##_1), t, QB_##P##_0)

constexpr float kInvPi = 0.318309886183790672f;
constexpr float kPiHi = 3.140625f;                    // 8 significant bits: n*kPiHi exact for |n| < 2^16
constexpr float kPiLo = 9.67653589793e-4f;            // pi - kPiHi
constexpr float kMagic = 12582912.0f;                 // 1.5 * 2^23: float add rounds to integer
constexpr float kPiO4 = 0.785398163397448310f;
constexpr float k3PiO4 = 2.35619449019234493f;

// cos(theta) * amp for theta of moderate size, sign folded through the parity of n.
__device__ __forceinline__ float amp_cos(float amp, float theta) {
    float t = fmaf(theta, kInvPi, kMagic);            // integer part lands in the low mantissa bits
    float n = t - kMagic;
    float r = fmaf(n, -kPiHi, theta);
    r = fmaf(n, -kPiLo, r);                           // r in [-pi/2, pi/2]
    float s = r * r;
    float c = QB_H5(CS, s);
    unsigned sign = __float_as_uint(t) << 31;         // (-1)^n
    return __uint_as_float(__float_as_uint(amp) ^ sign) * c;
}

template <bool WANT_J1>
__device__ __forceinline__ void bessel_small(float x, float& omj0, float& j1) {
    float z = x * x;
    omj0 = z * QB_H5(S0, z);
    if (WANT_J1) j1 = x * QB_H5(S1, z);
}

// MUFU.RSQ without the subnormal-input fix-up rsqrtf() carries (x > QB_X_SPLIT here);
// same 2-ulp unit, not a reduced-precision intrinsic.
__device__ __forceinline__ float rsqrt_pos(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

template <bool WANT_J1>
__device__ __forceinline__ void bessel_big(float x, float& omj0, float& j1) {
    float r = rsqrt_pos(x);
    float q = r * r;
    float w = q * q;
    float a0 = r * QB_H4(A0, w);
    float th0 = fmaf(q, QB_H4(F0, w), x - kPiO4);
    omj0 = 1.0f - amp_cos(a0, th0);
    if (WANT_J1) {
        float a1 = r * QB_H4(A1, w);
        float th1 = fmaf(q, QB_H4(F1, w), x - k3PiO4);
        j1 = amp_cos(a1, th1);
    }
}

// x >= 0.  Per-lane branch: uniform within a warp pass for all but the pass that
// straddles QB_X_SPLIT (lanes hold consecutive quadrature nodes, x is monotone in lane).
template <bool WANT_J1>
__device__ __forceinline__ void bessel_pair(float x, float& omj0, float& j1) {
    if (x <= QB_X_SPLIT) bessel_small<WANT_J1>(x, omj0, j1);
    else bessel_big<WANT_J1>(x, omj0, j1);
}

}  // namespace qb
