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
            bessel_pair<BWD>(a[j] * u, f0, f1);
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
        v.kappa = P.kappa;
    } else {
        v.G = 0.f;
        v.dG_doef = 0.f;
        v.kappa = 1.0f;
    }
    v.bw = v.kappa * dbv;
    return v;
}

// Tissue signal of the log-linear model (signals.py:194-207) for one tau, with partials.
__device__ __forceinline__ void loglinear_tissue(const QboldParams& P, const VoxelPhys& v, float tau,
                                                 float& st, float& dst_doef, float& dst_ddbv) {
    const float tc = 1.0f / v.dw;
    const float r2p = v.dw * v.dbv;
    const float rt = r2p * tau;
    if (fabsf(tau) < tc) {
        const float e = -(0.3f * (rt * rt)) / v.dbv;
        st = P.e_tissue * expf(e);
        // e = -0.3 tau^2 dw^2 dbv
        const float de_drt = -(0.3f * 2.0f * rt) / v.dbv;
        dst_doef = st * de_drt * tau * v.dbv * v.dw_k;
        dst_ddbv = st * (de_drt * tau * v.dw + (0.3f * (rt * rt)) / (v.dbv * v.dbv));
    } else {
        st = P.e_tissue * expf(v.dbv - rt);
        dst_doef = st * (-tau * v.dbv * v.dw_k);
        dst_ddbv = st * (1.0f - tau * v.dw);
    }
}

// Signal of one (voxel, tau) from the tissue integral; lane-local (signals.py:98-114,169-172).
struct TauSignal {
    float S;         // mixed signal
    float dS_doef;   // partials of S (BWD)
    float dS_ddbv;
};

template <bool BWD>
__device__ __forceinline__ TauSignal tau_signal(const QboldParams& P, const VoxelPhys& v, float tau,
                                                float blood_b, float I, float D) {
    TauSignal r;
    float st, dst_doef = 0.f, dst_ddbv = 0.f;
    if (P.full_model) {
        st = expf(-v.dbv * I) * P.e_tissue;
        if (BWD) {
            // dI/doef = D * da/doef, a = 1.5 |tau| dw
            dst_doef = -v.dbv * st * (D * (1.5f * fabsf(tau)) * v.dw_k);
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
    } else {
        r.dS_doef = r.dS_ddbv = 0.f;
    }
    return r;
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

}  // namespace qb
