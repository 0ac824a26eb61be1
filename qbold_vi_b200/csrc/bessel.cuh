// Device-side FP32 evaluation of the pair (1 - J0(x), J1(x)) for x >= 0.
//
// The qBOLD tissue integrand (reference signals.py:169-171) needs 1 - J0, never J0 itself, and
// TensorFlow's gradient of bessel_j0 is -bessel_j1.  Both are evaluated here without fast-math
// intrinsics (no __cosf/__expf/__fdividef), in three ranges whose polynomial coefficients are
// immediates (tools/fit_bessel.py -> bessel_coef.h):
//
//   x <= 3      : 1 - J0 = z*S0(z), J1 = x*S1(z), z = x*x.  No cancellation for small x, where the
//                 quadrature weights ~ 1/u^2 are largest (relative error 1.6e-7).
//   3 < x <= 9  : degree-12 polynomials in t = x - 5.5 for 1 - J0 and J1 (25 FMA-pipe instructions for
//                 the pair; the modulus/phase form costs twice that).  Fitted on [2, 9]: a warp pass whose
//                 arguments straddle x = 3 can run this kernel on all lanes instead of diverging.
//   x > 9       : valid from x = 6.5 (same reason).  Modulus/phase form  J_n = rsqrt(x) A_n(w) cos(x - (2n+1)pi/4 + q F_n(w)),
//                 q = 1/x = rsqrt(x)^2, w = q^2, one MUFU.RSQ for the pair; the cosine is a polynomial
//                 in r^2 after a two-constant Cody-Waite reduction mod pi (n*PI_HI exact for |x| < 1e5).
//
// Max abs error vs float64 (float32 FMA arithmetic): 2.0e-7 / 2.2e-7 (small), 4.4e-7 / 3.4e-7 (mid, on [2, 9]),
// 4.9e-7 / 3.9e-7 (big, 6.5 <= x <= 40; the floor there is the float32 rounding of the phase, which the
// reference's Cephes kernels share).  Error budget: the signal tolerance 1e-5 allows ~1e-5 absolute on
// J0 at these nodes (their Simpson weights sum to < 5).
#pragma once
#include "bessel_coef.h"

namespace qb {

// Horner with immediate coefficients (P::c(i) folds to a literal after unrolling).
template <class P>
__device__ __forceinline__ float horner(float t) {
    float acc = P::c(P::N - 1);
#pragma unroll
    for (int i = P::N - 2; i >= 0; --i) acc = fmaf(acc, t, P::c(i));
    return acc;
}

constexpr float kInvPi = 0.318309886183790672f;
constexpr float kPiHi = 3.140625f;                    // 8 significant bits: n*kPiHi is exact
constexpr float kPiLo = 9.67653589793e-4f;            // pi - kPiHi
constexpr float kMagic = 12582912.0f;                 // 1.5 * 2^23: float add rounds to integer
constexpr float kPiO4 = 0.785398163397448310f;
constexpr float k3PiO4 = 2.35619449019234493f;

// amp * cos(theta); the sign (-1)^n is folded into amp through the parity bit of n.
__device__ __forceinline__ float amp_cos(float amp, float theta) {
    float t = fmaf(theta, kInvPi, kMagic);            // integer part lands in the low mantissa bits
    float n = t - kMagic;
    float r = fmaf(n, -kPiHi, theta);
    r = fmaf(n, -kPiLo, r);                           // r in [-pi/2, pi/2]
    float c = horner<coef::CS>(r * r);
    unsigned sign = __float_as_uint(t) << 31;         // (-1)^n
    return __uint_as_float(__float_as_uint(amp) ^ sign) * c;
}

// MUFU.RSQ without the subnormal-input fix-up rsqrtf() carries (x > 9 here); same 2-ulp unit.
#ifndef QB_HOST_EMU
__device__ __forceinline__ float rsqrt_pos(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
#else   // tests/host_emu: this header compiled for the host, every PTX statement replaced by its IEEE meaning
__device__ __forceinline__ float rsqrt_pos(float x) { return 1.0f / sqrtf(x); }
#endif

template <bool WANT_J1>
__device__ __forceinline__ void bessel_small(float x, float& omj0, float& j1) {
    float z = x * x;
    omj0 = z * horner<coef::S0>(z);
    if (WANT_J1) j1 = x * horner<coef::S1>(z);
}

template <bool WANT_J1>
__device__ __forceinline__ void bessel_mid(float x, float& omj0, float& j1) {
    float t = x - coef::kXC;
    omj0 = horner<coef::M0>(t);
    if (WANT_J1) j1 = horner<coef::M1>(t);
}

template <bool WANT_J1>
__device__ __forceinline__ void bessel_big(float x, float& omj0, float& j1) {
    float r = rsqrt_pos(x);
    float q = r * r;
    float w = q * q;
    float a0 = r * horner<coef::A0>(w);
    float th0 = fmaf(q, horner<coef::F0>(w), x - kPiO4);
    omj0 = 1.0f - amp_cos(a0, th0);
    if (WANT_J1) {
        float a1 = r * horner<coef::A1>(w);
        float th1 = fmaf(q, horner<coef::F1>(w), x - k3PiO4);
        j1 = amp_cos(a1, th1);
    }
}


// ---------------------------------------------------------------------------------------------
// Packed FP32 (sm_100a fma/mul/add.rn.f32x2 -> SASS FFMA2 / FMUL2 / FADD2): one instruction evaluates the same
// Horner step for TWO quadrature entries held in an aligned register pair.  The polynomial coefficients stay
// 32-bit immediates (FFMA2 broadcasts them), pack/unpack of the two halves is free (they are the two registers
// of the pair), so the MUFU.RSQ and the sign-bit fix-up of the large-argument form stay scalar on each half.
// Same IEEE round-to-nearest arithmetic per half as the scalar kernels above.
typedef unsigned long long f32x2;

#ifdef QB_HOST_EMU
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
    return (f32x2)__float_as_uint(lo) | ((f32x2)__float_as_uint(hi) << 32);
}
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) {
    lo = __uint_as_float((unsigned)(v & 0xffffffffull));
    hi = __uint_as_float((unsigned)(v >> 32));
}
#define QB_EMU_OP2(NAME, EXPR)                                                        \
    __device__ __forceinline__ f32x2 NAME {                                           \
        float a0, a1, b0, b1, c0 = 0.f, c1 = 0.f;                                     \
        upk2(a, a0, a1);                                                              \
        upk2(b, b0, b1);                                                              \
        EXPR;                                                                         \
    }
QB_EMU_OP2(fma2(f32x2 a, f32x2 b, f32x2 c), upk2(c, c0, c1); return pk2(fmaf(a0, b0, c0), fmaf(a1, b1, c1)))
QB_EMU_OP2(mul2(f32x2 a, f32x2 b), (void)c0; (void)c1; return pk2(a0 * b0, a1 * b1))
QB_EMU_OP2(add2(f32x2 a, f32x2 b), (void)c0; (void)c1; return pk2(a0 + b0, a1 + b1))
#undef QB_EMU_OP2
#else
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
#endif   // QB_HOST_EMU
__device__ __forceinline__ f32x2 pk1(float v) { return pk2(v, v); }
__device__ __forceinline__ float hsum2(f32x2 v) {
    float lo, hi;
    upk2(v, lo, hi);
    return lo + hi;
}

template <class P>
__device__ __forceinline__ f32x2 horner2(f32x2 t) {
    f32x2 acc = pk1(P::c(P::N - 1));
#pragma unroll
    for (int i = P::N - 2; i >= 0; --i) acc = fma2(acc, t, pk1(P::c(i)));
    return acc;
}

// amp * cos(theta) on both halves (see amp_cos)
__device__ __forceinline__ f32x2 amp_cos2(f32x2 amp, f32x2 theta) {
    const f32x2 t = fma2(theta, pk1(kInvPi), pk1(kMagic));
    const f32x2 n = add2(t, pk1(-kMagic));
    f32x2 r = fma2(n, pk1(-kPiHi), theta);
    r = fma2(n, pk1(-kPiLo), r);
    const f32x2 c = horner2<coef::CS>(mul2(r, r));
    float t0, t1, a0, a1;
    upk2(t, t0, t1);
    upk2(amp, a0, a1);
    a0 = __uint_as_float(__float_as_uint(a0) ^ (__float_as_uint(t0) << 31));
    a1 = __uint_as_float(__float_as_uint(a1) ^ (__float_as_uint(t1) << 31));
    return mul2(pk2(a0, a1), c);
}

// Packed accumulation steps of the scheduled quadrature: x = A*m (both halves), w = Simpson weights.
//   small:  accI += w z S0(z),  accS += w z S1(z)        (accS is scaled by 1/A at the flush: w x J1 / A = w m J1)
//   mid/big: accI += w (1 - J0),  accB += (w m) J1
template <bool BWD>
__device__ __forceinline__ void acc_small2(f32x2 x, f32x2 w, f32x2& accI, f32x2& accS) {
    const f32x2 z = mul2(x, x);
    const f32x2 wz = mul2(w, z);
    accI = fma2(wz, horner2<coef::S0>(z), accI);
    if (BWD) accS = fma2(wz, horner2<coef::S1>(z), accS);
}

template <bool BWD>
__device__ __forceinline__ void acc_mid2(f32x2 x, f32x2 m, f32x2 w, f32x2& accI, f32x2& accB) {
    const f32x2 t = add2(x, pk1(-coef::kXC));
    accI = fma2(w, horner2<coef::M0>(t), accI);
    if (BWD) accB = fma2(mul2(w, m), horner2<coef::M1>(t), accB);
}

template <bool BWD>
__device__ __forceinline__ void acc_big2(f32x2 x, f32x2 m, f32x2 w, f32x2& accI, f32x2& accB) {
    float x0, x1;
    upk2(x, x0, x1);
    const f32x2 r = pk2(rsqrt_pos(x0), rsqrt_pos(x1));
    const f32x2 q = mul2(r, r);
    const f32x2 v = mul2(q, q);
    const f32x2 a0 = mul2(r, horner2<coef::A0>(v));
    const f32x2 th0 = fma2(q, horner2<coef::F0>(v), add2(x, pk1(-kPiO4)));
    const f32x2 omj0 = fma2(amp_cos2(a0, th0), pk1(-1.0f), pk1(1.0f));
    accI = fma2(w, omj0, accI);
    if (BWD) {
        const f32x2 a1 = mul2(r, horner2<coef::A1>(v));
        const f32x2 th1 = fma2(q, horner2<coef::F1>(v), add2(x, pk1(-k3PiO4)));
        accB = fma2(mul2(w, m), amp_cos2(a1, th1), accB);
    }
}

// x >= 0.  Per-lane branches: uniform within a warp pass except for the (at most two) passes per
// column that straddle a range boundary (lanes hold consecutive quadrature nodes, x is monotone in lane).
template <bool WANT_J1>
__device__ __forceinline__ void bessel_pair(float x, float& omj0, float& j1) {
    if (x <= coef::kX1) bessel_small<WANT_J1>(x, omj0, j1);
    else if (x <= coef::kX2) bessel_mid<WANT_J1>(x, omj0, j1);
    else bessel_big<WANT_J1>(x, omj0, j1);
}

}  // namespace qb
