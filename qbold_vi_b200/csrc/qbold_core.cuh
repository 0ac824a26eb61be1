// Warp-cooperative building blocks of the qBOLD forward model (device side).
//
// Mapping: one warp owns one voxel at a time.  The 128 live Simpson
// nodes are spread over the 32 lanes in 4 passes of 32 consecutive nodes, so the Bessel
// argument x = a_j * u_k is monotone in the lane index and the small/large-argument branch
// is warp-uniform for every pass except the one straddling the split.  Each lane keeps
// one partial sum per tau column; a butterfly transpose-reduce turns the 8 x 32 partials
// into 8 totals with 9 shuffles (instead of 40).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/qbold.h"
#include "bessel.cuh"

namespace qb {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kColGroup = 8;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// On return the lane holds the warp-wide total of v[4*b4 + 2*b3 + b2] (b_i = bit i of lane).
__device__ __forceinline__ float butterfly8(const float (&v)[8], int lane) {
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
    float w4[4], w2[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float keep = b4 ? v[i + 4] : v[i];
        float send = b4 ? v[i] : v[i + 4];
        w4[i] = keep + __shfl_xor_sync(kFull, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        float keep = b3 ? w4[i + 2] : w4[i];
        float send = b3 ? w4[i] : w4[i + 2];
        w2[i] = keep + __shfl_xor_sync(kFull, send, 8);
    }
    float keep = b2 ? w2[1] : w2[0];
    float send = b2 ? w2[0] : w2[1];
    float w1 = keep + __shfl_xor_sync(kFull, send, 4);
    w1 += __shfl_xor_sync(kFull, w1, 2);
    w1 += __shfl_xor_sync(kFull, w1, 1);
    return w1;
}

__device__ __forceinline__ int butterfly8_src_lane(int col) {
    return (((col >> 2) & 1) << 4) | (((col >> 1) & 1) << 3) | ((col & 1) << 2);
}

// Quadrature tables of one CTA in shared memory: [u | c | d] x 128 live nodes.
struct QuadSmem {
    float u[128], c[128], d[128];
};

__device__ __forceinline__ void load_quad_tables(const QboldParams& P, QuadSmem& s) {
    for (int i = threadIdx.x; i < 128; i += blockDim.x) {
        s.u[i] = P.qu[i];
        s.c[i] = P.qc[i];
        s.d[i] = P.qd[i];
    }
}

// Tissue integrals of one voxel (reference signals.py:159-185).
//   I_j  = sum_k qc[k] * (1 - J0(a_j u_k)),  a_j = 1.5 * (|tau_j| * dw)
//   D_j  = sum_k qd[k] * J1(a_j u_k) = dI_j/da_j      (BWD only)
// Lane t < n_tau receives (I, D) of its own tau (0 for tau == 0); my_col = col_of_tau[t] or -1.
// atau0 holds the first group of (up to 8) distinct |tau| values, hoisted out of the voxel
// loop by the caller; further groups (n_cols > 8, e.g. the 24-tau grid) are read from P.
struct TauCols {
    float atau[kColGroup];
};

__device__ __forceinline__ TauCols load_tau_cols(const QboldParams& P, int g) {
    TauCols t;
#pragma unroll
    for (int j = 0; j < kColGroup; ++j) {
        const int col = g * kColGroup + j;
        t.atau[j] = (col < P.n_cols) ? P.abs_tau[col] : 0.f;
    }
    return t;
}

template <bool BWD>
__device__ __forceinline__ void tissue_group(const QuadSmem& s, const TauCols& tc, float dw, int lane,
                                             float& ti, float& td) {
    float a[kColGroup], accI[kColGroup], accD[kColGroup];
#pragma unroll
    for (int j = 0; j < kColGroup; ++j) {
        a[j] = 1.5f * (tc.atau[j] * dw);
        accI[j] = accD[j] = 0.f;
    }
#pragma unroll 1
    for (int p = 0; p < 4; ++p) {
        const float u = s.u[p * 32 + lane];
        const float c = s.c[p * 32 + lane];
        const float d = s.d[p * 32 + lane];
#pragma unroll
        for (int j = 0; j < kColGroup; ++j) {
            float f0, f1 = 0.f;
#ifdef QB_SIGNED_OEF
            bessel_pair<BWD>(fabsf(a[j] * u), f0, f1);              // see tissue_sched: even / odd continuation
            if (a[j] < 0.f) f1 = -f1;
#else
            bessel_pair<BWD>(a[j] * u, f0, f1);
#endif
            accI[j] = fmaf(c, f0, accI[j]);
            if (BWD) accD[j] = fmaf(d, f1, accD[j]);
        }
    }
    ti = butterfly8(accI, lane);
    td = 0.f;
    if (BWD) td = butterfly8(accD, lane);
}

// MULTI = false: at most 8 distinct |tau| columns (the 11-tau optimal.yaml grid): one group, nothing but
// tc0 is read.  MULTI = true (e.g. the 24-tau grid, 16 columns): further groups are read from P.
template <bool BWD, bool MULTI>
__device__ __forceinline__ void tissue_integrals(const QboldParams& P, const QuadSmem& s, const TauCols& tc0,
                                                 float dw, int lane, int my_col, float& I_out, float& D_out) {
    if (!MULTI) {
        float ti, td;
        tissue_group<BWD>(s, tc0, dw, lane, ti, td);
        const int src = butterfly8_src_lane(my_col & 7);
        const float vi = __shfl_sync(kFull, ti, src);
        const float vd = BWD ? __shfl_sync(kFull, td, src) : 0.f;
        I_out = (my_col >= 0) ? vi : 0.f;
        D_out = (my_col >= 0) ? vd : 0.f;
        return;
    }
    I_out = 0.f;
    D_out = 0.f;
    for (int g = 0; g * kColGroup < P.n_cols; ++g) {
        float ti, td;
        const TauCols tc = load_tau_cols(P, g);
        tissue_group<BWD>(s, tc, dw, lane, ti, td);
        const int c = my_col - g * kColGroup;
        const bool mine = (c >= 0) && (c < kColGroup);
        const int src = butterfly8_src_lane(c & 7);
        const float vi = __shfl_sync(kFull, ti, src);
        const float vd = BWD ? __shfl_sync(kFull, td, src) : 0.f;
        if (mine) {
            I_out = vi;
            D_out = vd;
        }
    }
}

// Per-voxel scalars shared by all lanes.
struct VoxelPhys {
    float oef, dbv;
    float dw;        // delta omega                     signals.py:187
    float dw_k;      // d dw / d oef
    float G;         // blood exponent factor            signals.py:239-241
    float dG_doef;
    float bw;        // blood weight                     signals.py:107/110
    float kappa;     // d bw / d dbv
    float hratio;    // (d dw / d hct) / (d dw / d oef) = oef / hct      (variable_hct only)
    float dG_dhct;   //                                                 (variable_hct only)
};

template <bool HCT>
__device__ __forceinline__ VoxelPhys voxel_phys(const QboldParams& P, float oef, float dbv, float hct) {
    VoxelPhys v;
    v.oef = oef;
    v.dbv = dbv;
    v.dw_k = HCT ? P.dw_k_nohct * hct : P.dw_k;
    v.dw = v.dw_k * oef;
    if (P.include_blood) {
        const float c0 = HCT ? ((4.0f / 45.0f) * hct) * (1.0f - hct) : P.blood_c0;
        const float base = P.blood_c1 * oef;
        const float g0 = c0 * (base * base);
        v.G = (P.blood_hg * g0) * P.blood_td2;
        v.dG_doef = ((P.blood_hg * (c0 * (2.0f * base * P.blood_c1))) * P.blood_td2);
        v.dG_dhct = HCT ? (P.blood_hg * (((4.0f / 45.0f) * (1.0f - 2.0f * hct)) * (base * base))) * P.blood_td2 : 0.f;
        v.kappa = P.kappa;
    } else {
        v.G = 0.f;
        v.dG_doef = 0.f;
        v.dG_dhct = 0.f;
        v.kappa = 1.0f;
    }
    v.hratio = (HCT && hct != 0.f) ? oef / hct : 0.f;
    v.bw = v.kappa * dbv;
    return v;
}

// Tissue signal of the log-linear model (signals.py:194-207) for one tau, with partials.
// Cold path (full_model = False goes through k_loglinear): kept out of line so it does not bloat the
// instruction footprint of the quadrature kernels.
__device__ __forceinline__ void loglinear_tissue(const QboldParams& P, const VoxelPhys& v, float tau,
                                                 float& st, float& dst_doef, float& dst_ddbv) {
    // One expf per element: the branch (signals.py:199-205) selects the exponent, not the exponential, so a warp
    // whose voxels straddle |tau| = 1/dw does not evaluate both.  1/dw and 1/dbv are per-voxel (hoisted by the
    // compiler out of the tau loop); -(0.3 rt^2)/dbv is evaluated as -(0.3 rt^2) * (1/dbv) (<= 1 ulp apart).
    const float tc = 1.0f / v.dw;
    const float inv_dbv = 1.0f / v.dbv;
    const float r2p = v.dw * v.dbv;
    const float rt = r2p * tau;
    const bool short_tau = fabsf(tau) < tc;
    const float q = 0.3f * (rt * rt);
    st = P.e_tissue * expf(short_tau ? -(q * inv_dbv) : v.dbv - rt);
    // short: e = -0.3 tau^2 dw^2 dbv
    const float de_drt = -(0.3f * 2.0f * rt) * inv_dbv;
    dst_doef = short_tau ? st * de_drt * tau * v.dbv * v.dw_k : st * (-tau * v.dbv * v.dw_k);
    dst_ddbv = short_tau ? st * (de_drt * tau * v.dw + q * (inv_dbv * inv_dbv)) : st * (1.0f - tau * v.dw);
}

// Signal of one (voxel, tau) from the tissue integral; lane-local (signals.py:98-114,169-172).
struct TauSignal {
    float S;         // mixed signal
    float dS_doef;   // partials of S (BWD)
    float dS_ddbv;
    float dS_dhct;   // variable_hct: the tissue term depends on Hct through dw only (dw = k hct oef), the blood
                     // term through G0 ~ hct (1 - hct)   (signals.py:142-144, 239)
};

// I = tissue integral of this tau, dI_doef = its derivative w.r.t. OEF (BWD).
template <bool BWD>
__device__ __forceinline__ TauSignal tau_signal(const QboldParams& P, const VoxelPhys& v, float tau,
                                                float blood_b, float I, float dI_doef) {
    TauSignal r;
    float st, dst_doef = 0.f, dst_ddbv = 0.f;
    if (P.full_model) {
        st = expf(-v.dbv * I) * P.e_tissue;
        if (BWD) {
            dst_doef = -v.dbv * st * dI_doef;
            dst_ddbv = -I * st;
        }
    } else {
        loglinear_tissue(P, v, tau, st, dst_doef, dst_ddbv);
    }
    float sb = 0.f, dsb_doef = 0.f;
    if (P.include_blood) {
        sb = P.e_blood * expf(-v.G * blood_b);
        if (BWD) dsb_doef = -blood_b * sb * v.dG_doef;
    }
    const float tw = 1.0f - v.bw;
    r.S = tw * st + v.bw * sb;
    if (BWD) {
        r.dS_doef = tw * dst_doef + v.bw * dsb_doef;
        r.dS_ddbv = tw * dst_ddbv + v.kappa * (sb - st);
        r.dS_dhct = (tw * dst_doef) * v.hratio + v.bw * (-blood_b * sb * v.dG_dhct);
    } else {
        r.dS_doef = r.dS_ddbv = r.dS_dhct = 0.f;
    }
    return r;
}

// Dynamic work distribution for kernels with uneven per-unit cost (a masked voxel costs nothing, a live one tens of
// microseconds; with a static warp stride that is a multiple of the volume's z extent whole warps only ever see
// voxels outside the mask).  Lane 0 takes the next unit from a device counter, the warp follows.
__device__ __forceinline__ int64_t next_unit(unsigned long long* counter, int lane) {
    unsigned long long v = 0;
    if (lane == 0) v = atomicAdd(counter, 1ull);
    return (int64_t)__shfl_sync(kFull, v, 0);
}

// FP32 artefact of the reference at quadrature node 0: TensorFlow evaluates
// 1 - j0f(x0) with the Cephes tiny-argument branch 1 - 0.25 x0^2, which rounds to exactly 1
// (so the node contributes 0) for every admissible voxel.  Reproduced literally so that
// out-of-domain inputs (x0 >= ~3.5e-4) still follow the reference.
__device__ __forceinline__ float node0_value(const QboldParams& P, float a) {
    const float x0 = a * P.qu[0];
    const float z = __fmul_rn(x0, x0);
    const float j = __fsub_rn(1.0f, __fmul_rn(0.25f, z));
    return P.node0_c * __fsub_rn(1.0f, j);
}

// ---------------------------------------------------------------------------------------------
// Scheduled path (n_cols <= 8): the static lane schedule of QboldParams::sched_* (built on the host).
// Every pass hands each lane one (column, node) pair with a similar Bessel argument x = A*m, so a
// whole phase of 4 passes is usually in ONE argument range and runs branch-free straight-line code;
// a lane accumulates a single column per phase and flushes into its private shared-memory slot
// [column][lane] at the phase end (first visit stores, later visits add -- no per-voxel clearing).
enum QuadPath { kSched = 0, kCols = 1, kColsMulti = 2 };

// Which argument ranges evaluate two schedule entries per instruction with packed FP32 (FFMA2).
#ifndef QB_PACK_SMALL
#define QB_PACK_SMALL 1
#endif
#ifndef QB_PACK_MID
#define QB_PACK_MID 1
#endif
#ifndef QB_PACK_BIG
#define QB_PACK_BIG 1
#endif

struct SchedSmem {
    // packed entries [phase][pair][lane] = (m_a, m_b, w_a, w_b): passes 2*pair and 2*pair+1 of the host schedule,
    // evaluated together with packed FP32 (one LDS.128 per lane feeds two Bessel-pair evaluations)
    float4 mw[QBOLD_SCHED_MAX_ENTRIES / 2];
    unsigned char col[QBOLD_SCHED_MAX_PHASES * 32];
    float slots[8 /*warps*/][2][8][32];                  // [warp][I|D][column][lane]
};

__device__ __forceinline__ void load_sched(const QboldParams& P, SchedSmem& s) {
    static_assert(QBOLD_SCHED_PHASE_LEN % 2 == 0, "packed evaluation pairs up the passes of a phase");
    const int n = P.sched_phases * QBOLD_SCHED_PHASE_LEN * 32;
    float* mw = reinterpret_cast<float*>(s.mw);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int lane = i & 31, pass = i >> 5, pair = pass >> 1, half = pass & 1;
        float* e = mw + (pair * 32 + lane) * 4;
        e[half] = P.sched_m[i];
        e[2 + half] = P.sched_w[i];
    }
    for (int i = threadIdx.x; i < P.sched_phases * 32; i += blockDim.x) s.col[i] = P.sched_col[i];
    float* z = &s.slots[0][0][0][0];
    for (int i = threadIdx.x; i < 8 * 2 * 8 * 32; i += blockDim.x) z[i] = 0.f;
}

// Shared-memory accesses of the scheduled path use explicit 32-bit shared addresses: one LDS/STS with a
// register base + immediate offset, no generic-to-shared address arithmetic inside the voxel loop.
#ifdef QB_HOST_EMU
// tests/host_emu: the same code on the host.  A "shared address" is the 32-bit offset of the object from the
// emulator's anchor (qb_emu::smem_anchor, same module), so the register-base + immediate-offset arithmetic of the
// scheduled path runs unchanged; the ld/st.shared statements become plain loads and stores.
__device__ __forceinline__ unsigned smem_addr(const void* p) { return qb_emu::to_shared(p); }
__device__ __forceinline__ float2 lds_f2(unsigned a) { return *static_cast<const float2*>(qb_emu::from_shared(a)); }
__device__ __forceinline__ float4 lds_f4(unsigned a) { return *static_cast<const float4*>(qb_emu::from_shared(a)); }
__device__ __forceinline__ float lds_f1(unsigned a) { return *static_cast<const float*>(qb_emu::from_shared(a)); }
__device__ __forceinline__ unsigned lds_u8(unsigned a) { return *static_cast<const unsigned char*>(qb_emu::from_shared(a)); }
__device__ __forceinline__ void sts_f1(unsigned a, float v) { *static_cast<float*>(qb_emu::from_shared(a)) = v; }
#else
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float2 lds_f2(unsigned a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ float4 lds_f4(unsigned a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ float lds_f1(unsigned a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ unsigned lds_u8(unsigned a) {
    unsigned v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_f1(unsigned a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
#endif   // QB_HOST_EMU

template <bool BWD>
__device__ __forceinline__ void acc_small(float x, float w, float& accI, float& accS) {
    const float z = x * x;
    const float wz = w * z;
    accI = fmaf(wz, horner<coef::S0>(z), accI);                  // w * (1 - J0) = w z S0(z)
    if (BWD) accS = fmaf(wz, horner<coef::S1>(z), accS);         // w * x J1 = w z S1(z)   (scaled by 1/A later)
}

template <bool BWD>
__device__ __forceinline__ void acc_mid(float x, float2 e, float& accI, float& accB) {
    float f0, j1 = 0.f;
    bessel_mid<BWD>(x, f0, j1);
    accI = fmaf(e.y, f0, accI);
    if (BWD) accB = fmaf(e.y * e.x, j1, accB);                   // w * m * J1
}

template <bool BWD>
__device__ __forceinline__ void acc_big(float x, float2 e, float& accI, float& accB) {
    float f0, j1 = 0.f;
    bessel_big<BWD>(x, f0, j1);
    accI = fmaf(e.y, f0, accI);
    if (BWD) accB = fmaf(e.y * e.x, j1, accB);
}

// Per-warp shared addresses of the scheduled path, computed once per kernel.
struct SchedAddr {
    unsigned mw;      // &mw[lane]
    unsigned col;     // &col[lane]
    unsigned slotI;   // &slots[warp][0][0][lane]
    unsigned slotD;   // &slots[warp][1][0][lane]
    unsigned redI;    // &slots[warp][0][lane >> 2][(lane & 3) * 8]
    unsigned redD;
};

// The addresses go through an empty asm so that the compiler keeps them in registers: left to itself it
// re-derives them from %tid / the CTA's shared window at every phase flush (~10 instructions per phase).
__device__ __forceinline__ unsigned pinned(unsigned a) {
    asm volatile("" : "+r"(a));
    return a;
}

__device__ __forceinline__ SchedAddr sched_addr(const SchedSmem& s, int lane, int warp_in_cta) {
    SchedAddr a;
    a.mw = pinned(smem_addr(&s.mw[lane]));
    a.col = pinned(smem_addr(&s.col[lane]));
    a.slotI = pinned(smem_addr(&s.slots[warp_in_cta][0][0][lane]));
    a.slotD = a.slotI + (unsigned)sizeof(float) * 8 * 32;
    a.redI = pinned(smem_addr(&s.slots[warp_in_cta][0][lane >> 2][(lane & 3) * 8]));
    a.redD = a.redI + (unsigned)sizeof(float) * 8 * 32;
    return a;
}

// Returns for lane t < n_tau: I = sum_{k>=1} c_k (1 - J0(a_t u_k)) and
// Dm = sum_{k>=1} c_k m_tk J1(a_t u_k)  (m = |tau_t|/tau_ref * u_k), so dI/dA = Dm with a_t u_k = A m.
// ph_lo / ph_hi: this lane's phase bounds (lane p < sched_phases holds phase p), hoisted by the caller.
// Every pass evaluates TWO schedule entries per lane with packed FP32 (FFMA2): the kernels are issue-bound, and the
// Horner chains of two entries cost one instruction slot per step instead of two.
template <bool BWD>
__device__ __forceinline__ void tissue_sched(int nph, const SchedAddr& sa, float A, int lane, float ph_lo, float ph_hi,
                                             int my_col, float& I_out, float& Dm_out) {
    constexpr int kPairBytes = 32 * 16, kPairs = QBOLD_SCHED_PHASE_LEN / 2, kPhaseBytes = kPairs * kPairBytes;
    const bool is_ph = lane < nph;
#ifdef QB_SIGNED_OEF
    // Negative OEF (outside the reference's callers): 1 - J0 is even and J1 odd in A, so the
    // quadrature runs on |A| and Dm takes the sign of A.  Off by default: verified in the host emulator only, to be
    // switched on together with a GPU measurement of the kernels it touches.
    const float a_sign = A < 0.f ? -1.0f : 1.0f;
    A = fabsf(A);
#endif
    const float lo = A * ph_lo, hi = A * ph_hi;
    // one kernel for the whole phase whenever its argument span fits the kernel's validity range
    // (small: x <= 3, mid: [2, 9], big: x >= 6.5 -- the ranges overlap on purpose, see bessel.cuh)
    const unsigned PS = __ballot_sync(kFull, is_ph && hi <= coef::kX1);
    const unsigned PM = __ballot_sync(kFull, is_ph && lo >= coef::kMidLo && hi <= coef::kX2);
    const unsigned PB = __ballot_sync(kFull, is_ph && lo >= coef::kBigLo);
    const float invA = A > 0.f ? 1.0f / A : 0.f;
    const f32x2 A2 = pk1(A);
    f32x2 accI = pk1(0.f), accS = pk1(0.f), accB = pk1(0.f);
    unsigned ea = sa.mw, ca = sa.col;
#pragma unroll 1
    for (int ph = 0; ph < nph; ++ph, ea += kPhaseBytes, ca += 32) {
        const unsigned bit = 1u << ph;
        if (PS & bit) {
#pragma unroll
            for (int c = 0; c < kPairs; ++c) {
                const float4 e = lds_f4(ea + c * kPairBytes);
                if (QB_PACK_SMALL) {
                    acc_small2<BWD>(mul2(A2, pk2(e.x, e.y)), pk2(e.z, e.w), accI, accS);
                } else {
                    float i0, i1, s0, s1;
                    upk2(accI, i0, i1);
                    upk2(accS, s0, s1);
                    acc_small<BWD>(A * e.x, e.z, i0, s0);
                    acc_small<BWD>(A * e.y, e.w, i1, s1);
                    accI = pk2(i0, i1);
                    accS = pk2(s0, s1);
                }
            }
        } else if (PM & bit) {
#pragma unroll
            for (int c = 0; c < kPairs; ++c) {
                const float4 e = lds_f4(ea + c * kPairBytes);
                if (QB_PACK_MID) {
                    const f32x2 m = pk2(e.x, e.y);
                    acc_mid2<BWD>(mul2(A2, m), m, pk2(e.z, e.w), accI, accB);
                } else {
                    float i0, i1, b0, b1;
                    upk2(accI, i0, i1);
                    upk2(accB, b0, b1);
                    acc_mid<BWD>(A * e.x, make_float2(e.x, e.z), i0, b0);
                    acc_mid<BWD>(A * e.y, make_float2(e.y, e.w), i1, b1);
                    accI = pk2(i0, i1);
                    accB = pk2(b0, b1);
                }
            }
        } else if (PB & bit) {
#pragma unroll
            for (int c = 0; c < kPairs; ++c) {
                const float4 e = lds_f4(ea + c * kPairBytes);
                if (QB_PACK_BIG) {
                    const f32x2 m = pk2(e.x, e.y);
                    acc_big2<BWD>(mul2(A2, m), m, pk2(e.z, e.w), accI, accB);
                } else {
                    float i0, i1, b0, b1;
                    upk2(accI, i0, i1);
                    upk2(accB, b0, b1);
                    acc_big<BWD>(A * e.x, make_float2(e.x, e.z), i0, b0);
                    acc_big<BWD>(A * e.y, make_float2(e.y, e.w), i1, b1);
                    accI = pk2(i0, i1);
                    accB = pk2(b0, b1);
                }
            }
        } else {
            // phase straddles a range boundary: decide per packed pass with warp votes (uniform branches); only
            // the pass that really straddles it takes the per-lane scalar path
#pragma unroll 1
            for (int c = 0; c < kPairs; ++c) {
                const float4 e = lds_f4(ea + c * kPairBytes);
                const float x0 = A * e.x, x1 = A * e.y;
                const float xlo = fminf(x0, x1), xhi = fmaxf(x0, x1);
                const f32x2 m = pk2(e.x, e.y), w = pk2(e.z, e.w), x = pk2(x0, x1);
                if (__all_sync(kFull, xhi <= coef::kX1)) {
                    acc_small2<BWD>(x, w, accI, accS);
                } else if (__all_sync(kFull, xlo >= coef::kMidLo && xhi <= coef::kX2)) {
                    acc_mid2<BWD>(x, m, w, accI, accB);
                } else if (__all_sync(kFull, xlo >= coef::kBigLo)) {
                    acc_big2<BWD>(x, m, w, accI, accB);
                } else {
                    float i0, i1, s0, s1, b0, b1;
                    upk2(accI, i0, i1);
                    upk2(accS, s0, s1);
                    upk2(accB, b0, b1);
                    if (x0 <= coef::kX1) acc_small<BWD>(x0, e.z, i0, s0);
                    else if (x0 <= coef::kX2) acc_mid<BWD>(x0, make_float2(e.x, e.z), i0, b0);
                    else acc_big<BWD>(x0, make_float2(e.x, e.z), i0, b0);
                    if (x1 <= coef::kX1) acc_small<BWD>(x1, e.w, i1, s1);
                    else if (x1 <= coef::kX2) acc_mid<BWD>(x1, make_float2(e.y, e.w), i1, b1);
                    else acc_big<BWD>(x1, make_float2(e.y, e.w), i1, b1);
                    accI = pk2(i0, i1);
                    accS = pk2(s0, s1);
                    accB = pk2(b0, b1);
                }
            }
        }
        // flush this phase's partial sums into the lane's private slot of its column
        const unsigned cc = lds_u8(ca);
        const unsigned off = (cc & 7u) << 7;                     // column * 32 floats
        const bool first = cc & 0x80u;
        float oldI = 0.f, oldD = 0.f;
        if (!first) {
            oldI = lds_f1(sa.slotI + off);
            if (BWD) oldD = lds_f1(sa.slotD + off);
        }
        sts_f1(sa.slotI + off, oldI + hsum2(accI));
        accI = pk1(0.f);
        if (BWD) {
            sts_f1(sa.slotD + off, oldD + fmaf(hsum2(accS), invA, hsum2(accB)));
            accS = accB = pk1(0.f);
        }
    }
    __syncwarp();
    // column totals: the 4 lanes of column j = lane >> 2 sum the 32 per-lane slots of that column
    const float4 a0 = lds_f4(sa.redI), a1 = lds_f4(sa.redI + 16);
    float ti = ((a0.x + a0.y) + (a0.z + a0.w)) + ((a1.x + a1.y) + (a1.z + a1.w));
    ti += __shfl_xor_sync(kFull, ti, 1);
    ti += __shfl_xor_sync(kFull, ti, 2);
    float td = 0.f;
    if (BWD) {
        const float4 b0 = lds_f4(sa.redD), b1 = lds_f4(sa.redD + 16);
        td = ((b0.x + b0.y) + (b0.z + b0.w)) + ((b1.x + b1.y) + (b1.z + b1.w));
        td += __shfl_xor_sync(kFull, td, 1);
        td += __shfl_xor_sync(kFull, td, 2);
    }
    __syncwarp();                                        // slots are rewritten by the next voxel's first flush
    const int src = (my_col & 7) << 2;
    const float vi = __shfl_sync(kFull, ti, src);
    const float vd = BWD ? __shfl_sync(kFull, td, src) : 0.f;
    I_out = (my_col >= 0) ? vi : 0.f;
    Dm_out = (my_col >= 0) ? vd : 0.f;
#ifdef QB_SIGNED_OEF
    Dm_out *= a_sign;
#endif
}

// One entry point for the three quadrature paths: returns (I, dI/dOEF) of this lane's tau.
struct QuadCtx {
    int lane, my_col, nph;
    float my_tau, ph_lo, ph_hi;
    float tau_ref15, node0_d;        // 1.5 * tau_ref;  qd[0] * 0.5 * u_0 * 1.5 |tau_t|  (node-0 derivative term)
    SchedAddr sa;
    TauCols tc0;
};

template <int PATH>
__device__ __forceinline__ QuadCtx make_quad_ctx(const QboldParams& P, const SchedSmem& ss, int lane, int my_col,
                                                 float my_tau) {
    QuadCtx c;
    c.lane = lane;
    c.my_col = my_col;
    c.my_tau = my_tau;
    c.nph = P.sched_phases;
    const bool ph = (PATH == kSched) && lane < P.sched_phases;
    c.ph_lo = ph ? P.sched_ph_min[lane & 15] : 0.f;
    c.ph_hi = ph ? P.sched_ph_max[lane & 15] : 0.f;
    c.tau_ref15 = 1.5f * P.tau_ref;
    c.node0_d = P.qd[0] * (0.5f * P.qu[0]) * (1.5f * fabsf(my_tau));
    if (PATH == kSched) c.sa = sched_addr(ss, lane, threadIdx.x >> 5);
    else c.tc0 = load_tau_cols(P, 0);
    return c;
}

template <bool BWD, int PATH>
__device__ __forceinline__ void tissue_eval(const QboldParams& P, const QuadSmem& qs, SchedSmem& ss, const QuadCtx& c,
                                            float dw, float dw_k, float& I, float& dI_doef) {
    if (PATH == kSched) {
        const float A = c.tau_ref15 * dw;
        float Dm;
        tissue_sched<BWD>(c.nph, c.sa, A, c.lane, c.ph_lo, c.ph_hi, c.my_col, I, Dm);
        if (BWD) {
            // dI/dOEF = Dm * dA/dOEF + node 0 (live in the derivative only; J1(x) = x/2 for x ~ 1e-4)
            const float a_t = 1.5f * (fabsf(c.my_tau) * dw);
            dI_doef = (Dm * c.tau_ref15 + c.node0_d * a_t) * dw_k;
        }
    } else {
        float D;
        tissue_integrals<BWD, PATH == kColsMulti>(P, qs, c.tc0, dw, c.lane, c.my_col, I, D);
        if (BWD) dI_doef = D * (1.5f * fabsf(c.my_tau)) * dw_k;
    }
    if (c.my_col >= 0) I += node0_value(P, 1.5f * (fabsf(c.my_tau) * dw));
}

}  // namespace qb
