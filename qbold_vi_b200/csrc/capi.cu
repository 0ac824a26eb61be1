// C-ABI plumbing: error reporting, parameter resolution (the host half of
// SignalGenerationLayer.__init__, reference signals.py:18-53), launch accounting and the
// FP32 FMA micro-benchmark used as a same-run roofline denominator.
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "launch.h"

namespace qb {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_check(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return QBOLD_OK;
    return fail(QBOLD_ECUDA, "%s: %s", what, cudaGetErrorString(e));
}

int after_launch(const char* kernel_name) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cuda_check(cudaGetLastError(), kernel_name);
}

int sm_count() {
    static int cached_dev = -1, cached = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != cached_dev) {
        cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
        cached_dev = dev;
        if (cached <= 0) cached = 148;
    }
    return cached;
}

// Ring size bounds how many launches may be in flight at once ACROSS streams before a counter is reused (launches on
// one stream are ordered with the memset that re-arms their counter, so a single stream can never collide).
constexpr unsigned kWorkCounters = 1024;
__device__ unsigned long long g_work_counters[kWorkCounters];

unsigned long long* next_work_counter(cudaStream_t stream) {
    static std::atomic<unsigned> turn{0};
    void* base = nullptr;
    if (cudaGetSymbolAddress(&base, g_work_counters) != cudaSuccess) return nullptr;     // per current device
    unsigned long long* c = static_cast<unsigned long long*>(base) + (turn.fetch_add(1) & (kWorkCounters - 1));
    if (cudaMemsetAsync(c, 0, sizeof(unsigned long long), stream) != cudaSuccess) return nullptr;
    return c;
}

// float32 exp of a Python-float argument, as tf.math.exp(<python float>) gives it:
// the argument is first converted to float32, the result is the float32 nearest exp().
static float expf_of(double x) { return (float)std::exp((double)(float)x); }

// Static lane schedule (see QboldParams::sched_* in include/qbold.h and DESIGN.md 3).
// Greedy n-way merge of the per-column node sequences by m = r_j*u_k decides, for each phase of
// QBOLD_SCHED_PHASE_LEN passes, how many lanes each column gets (largest-remainder rounding of its
// share of the next 4*32 entries); lanes are grouped by column, each column consumes its nodes in
// order.  Lagging columns get more lanes in the next phase, so the m-spread inside a pass stays small.
static void build_schedule(QboldParams& P) {
    P.sched_phases = 0;
    const int nc = P.n_cols, L = 32, PL = QBOLD_SCHED_PHASE_LEN, last = QBOLD_NQ - 2;   // live nodes 1..127
    if (nc < 1 || nc > 8) return;
    double tref = 0.0;
    for (int j = 0; j < nc; ++j) tref = std::fmax(tref, (double)P.abs_tau[j]);
    P.tau_ref = (float)tref;
    double r[8];
    int nxt[8];
    bool seen[32][8] = {};
    int remaining = nc * last;
    for (int j = 0; j < nc; ++j) {
        r[j] = (double)P.abs_tau[j] / tref;
        nxt[j] = 1;
    }
    int phase = 0;
    while (remaining > 0 && phase < QBOLD_SCHED_MAX_PHASES) {
        int tmp[8], cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, total = 0;
        for (int j = 0; j < nc; ++j) tmp[j] = nxt[j];
        for (int i = 0; i < PL * L; ++i) {
            int best = -1;
            double bm = 0.0;
            for (int j = 0; j < nc; ++j)
                if (tmp[j] <= last) {
                    const double m = r[j] * (double)P.qu[tmp[j]];
                    if (best < 0 || m < bm) {
                        best = j;
                        bm = m;
                    }
                }
            if (best < 0) break;
            ++tmp[best];
            ++cnt[best];
            ++total;
        }
        int n[8], sum = 0;
        double rem[8];
        for (int j = 0; j < nc; ++j) {
            const double q = (double)cnt[j] * L / (double)total;
            n[j] = (int)std::floor(q);
            rem[j] = cnt[j] > 0 ? q - n[j] : -1.0;
            sum += n[j];
        }
        while (sum < L) {
            int best = 0;
            for (int j = 1; j < nc; ++j)
                if (rem[j] > rem[best]) best = j;
            ++n[best];
            rem[best] = -1.0;
            ++sum;
        }
        int lane_col[32], lane = 0;
        for (int j = 0; j < nc; ++j)
            for (int i = 0; i < n[j] && lane < L; ++i) lane_col[lane++] = j;
        for (; lane < L; ++lane) lane_col[lane] = 0;
        double pmin = 1e300, pmax = 0.0;
        for (int c = 0; c < PL; ++c)
            for (int l = 0; l < L; ++l) {
                const int j = lane_col[l], e = (phase * PL + c) * L + l;
                if (nxt[j] <= last) {
                    const double m = r[j] * (double)P.qu[nxt[j]];
                    P.sched_m[e] = (float)m;
                    P.sched_w[e] = P.qc[nxt[j]];
                    pmin = std::fmin(pmin, m);
                    pmax = std::fmax(pmax, m);
                    ++nxt[j];
                    --remaining;
                } else {
                    P.sched_m[e] = 0.f;
                    P.sched_w[e] = 0.f;
                }
            }
        for (int l = 0; l < L; ++l) {
            const int j = lane_col[l];
            P.sched_col[phase * L + l] = (uint8_t)(j | (seen[l][j] ? 0 : 0x80));
            seen[l][j] = true;
        }
        // idle entries (weight 0) take the phase's largest m, so they sit in the same argument range as the
        // live ones (m = 0 would feed rsqrt(0) to the large-argument branch)
        for (int c = 0; c < PL; ++c)
            for (int l = 0; l < L; ++l) {
                const int e = (phase * PL + c) * L + l;
                if (P.sched_w[e] == 0.f) P.sched_m[e] = (float)pmax;
            }
        P.sched_ph_min[phase] = (float)(pmin > 1e200 ? 0.0 : pmin * (1.0 - 1e-6));
        P.sched_ph_max[phase] = (float)(pmax * (1.0 + 1e-6));
        ++phase;
    }
    if (remaining == 0) P.sched_phases = phase;   // else: keep 0 -> column-major fallback
}

}  // namespace qb

using namespace qb;

extern "C" int qbold_abi_version(void) { return QBOLD_ABI_VERSION; }
extern "C" int qbold_params_sizeof(void) { return (int)sizeof(QboldParams); }
extern "C" const char* qbold_last_error(void) { return g_err; }
extern "C" int64_t qbold_launch_count(void) { return g_launches.load(); }

extern "C" int qbold_params_init(QboldParams* out, const QboldPhysics* ph, const float* taus, int32_t n_tau,
                                 int32_t full_model, int32_t include_blood) {
    if (!out || !ph || !taus) return fail(QBOLD_EINVAL, "qbold_params_init: null pointer");
    if (n_tau < 1 || n_tau > QBOLD_MAX_TAU)
        return fail(QBOLD_EINVAL, "qbold_params_init: n_tau=%d outside [1,%d]", n_tau, QBOLD_MAX_TAU);
    std::memset(out, 0, sizeof(*out));
    QboldParams& P = *out;
    P.abi_version = QBOLD_ABI_VERSION;
    P.n_tau = n_tau;
    P.full_model = full_model ? 1 : 0;
    P.include_blood = include_blood ? 1 : 0;

    // ---- delta omega constant, signals.py:142-144: Python-double product, then * float32 tensor
    const double k_nohct = (4.0 / 3.0) * M_PI * ph->gamma * ph->b0 * ph->dchi;
    P.dw_k_nohct = (float)k_nohct;
    P.dw_k = (float)(k_nohct * ph->hct);
    P.hct = (float)ph->hct;
    P.e_tissue = expf_of(-ph->te * ph->r2t);                                   // signals.py:172

    // ---- blood compartment, signals.py:102-107 and :233-247
    const float e1 = expf_of(-(ph->tr - ph->ti) / ph->t1b);
    const float e2 = expf_of(-ph->ti / ph->t1b);
    const float m_bld = 1.0f - (2.0f - e1) * e2;
    P.kappa = m_bld * 0.775f;
    const double r2b = 1.0 / 0.189;
    const double td = ((2.6 * 2.6) / 2.0) * 1e-3;
    P.e_blood = expf_of(-r2b * ph->te);
    P.blood_c0 = (float)((4.0 / 45.0) * ph->hct * (1.0 - ph->hct));
    P.blood_c1 = (float)(4.0 * M_PI * ph->b0 * ph->dchi);
    P.blood_hg = (float)(0.5 * (ph->gamma * ph->gamma));
    P.blood_td2 = (float)(td * td);
    {
        const float tef = (float)ph->te, tdf = (float)td;
        float t0 = (float)(ph->te / td) + sqrtf((float)(0.25 + (ph->te / td)));
        t0 = t0 + 1.5f;
        for (int t = 0; t < n_tau; ++t) {
            const float s1 = 2.0f * sqrtf(0.25f + (tef + taus[t]) / tdf);
            const float s2 = 2.0f * sqrtf(0.25f + (tef - taus[t]) / tdf);
            P.blood_b[t] = (t0 - s1) - s2;
        }
    }

    // ---- tau grid and dedup of |tau|.  J0 is even, so tau and -tau share one quadrature column.
    // Magnitudes are matched from the ACTUAL tau array with a 4-ulp relative tolerance: depending on
    // the TensorFlow release tf.range yields start + i*delta or an accumulated sum, and e.g.
    // 3*0.008f - 0.016f differs from 0.008f by one ulp.  Sharing the column
    // perturbs the Bessel argument by <= 5e-7 relative, two orders below the 1e-5 signal tolerance.
    for (int t = 0; t < n_tau; ++t) {
        P.tau[t] = taus[t];
        const float at = fabsf(taus[t]);
        int col = -1;
        if (at != 0.0f) {
            for (int c = 0; c < P.n_cols && col < 0; ++c)
                if (fabsf(P.abs_tau[c] - at) <= 4.8e-7f * at) col = c;
            if (col < 0) {
                col = P.n_cols++;
                P.abs_tau[col] = at;
            }
        }
        P.col_of_tau[t] = col;
    }

    // ---- normalised SNR table, signals.py:117-121
    if (n_tau == 11) {
        static const float ns[11] = {0.985f, 1.00f, 1.01f, 1.f, 0.97f, 0.95f, 0.93f, 0.90f, 0.86f, 0.83f, 0.79f};
        for (int t = 0; t < 11; ++t) P.norm_snr[t] = ns[t];
    } else if (n_tau == 24) {
        for (int t = 0; t < 24; ++t) P.norm_snr[t] = (float)(1.0 - std::fabs(-0.028 + 0.004 * t) * 3.0);
    }

    // ---- quadrature table, signals.py:166-185 (float32 arithmetic in the reference's order)
    const float a = 1e-5f, b = 1.0f;
    const float delta = (b - a) / 128.0f;
    float u[QBOLD_NQ];
    for (int i = 0; i < QBOLD_NQ; ++i) u[i] = a + delta * (float)i;
    u[0] = a;
    u[QBOLD_NQ - 1] = b;
    const float h = (u[2] - u[0]) / 2.0f;
    const float h3 = h / 3.0f;
    for (int i = 0; i < QBOLD_NQ; ++i) {
        const float wk = (i == 0 || i == QBOLD_NQ - 1) ? 1.0f : ((i & 1) ? 4.0f : 2.0f);
        const float A = (2.0f + u[i]) * sqrtf(1.0f - u[i]);
        const float Dn = 3.0f * (u[i] * u[i]);
        const float c = (wk * h3) * (A / Dn);
        P.qu[i] = u[i];
        P.qc[i] = c;
        P.qd[i] = c * u[i];
    }
    P.node0_c = P.qc[0];
    P.qc[0] = 0.0f;   // node 0: exactly 0 in the reference's float32 value (handled by node0_value())

    build_schedule(P);

    // ---- default likelihood: optimal.yaml (Gaussian, single-image normalisation)
    P.se_idx = 0;
    P.student_t_df = 0.f;
    return QBOLD_OK;
}

extern "C" int qbold_params_set_likelihood(QboldParams* p, const QboldLikelihood* lik) {
    if (!p || !lik) return fail(QBOLD_EINVAL, "qbold_params_set_likelihood: null pointer");
    if (lik->se_idx < 0 || lik->se_idx >= p->n_tau)
        return fail(QBOLD_EINVAL, "se_idx=%d outside the tau grid (n_tau=%d)", lik->se_idx, p->n_tau);
    if (lik->multi_image_normalisation && (lik->se_idx < 1 || lik->se_idx + 1 >= p->n_tau))
        return fail(QBOLD_EINVAL, "multi_image_normalisation needs se_idx-1..se_idx+1 inside the tau grid");
    p->se_idx = lik->se_idx;
    p->multi_image_normalisation = lik->multi_image_normalisation ? 1 : 0;
    p->predict_log_data = lik->predict_log_data ? 1 : 0;
    const double df = lik->student_t_df;
    if (df > 0.0 && df < 50.0) {                                               // model.py:557
        p->student_t_df = (float)df;
        p->student_t_logc = (float)(std::lgamma(0.5 * (df + 1.0)) - std::lgamma(0.5 * df) - 0.5 * std::log(df * M_PI));
    } else {
        p->student_t_df = 0.f;
        p->student_t_logc = 0.f;
    }
    return QBOLD_OK;
}

// ---------------------------------------------------------------- FMA micro-benchmark
namespace qb {
__global__ void __launch_bounds__(256) k_fma_peak(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f;
    float x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}
}  // namespace qb

extern "C" int qbold_fma_peak(int32_t iters, double* tflops) {
    if (!tflops || iters < 1) return fail(QBOLD_EINVAL, "qbold_fma_peak: bad argument");
    const int blocks = sm_count() * 8, threads = 256;
    float* out = nullptr;
    int rc = cuda_check(cudaMalloc(&out, sizeof(float) * blocks * threads), "cudaMalloc");
    if (rc) return rc;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_fma_peak<<<blocks, threads>>>(out, 64, 0.999f, 1e-3f);   // warm-up
    after_launch("k_fma_peak");
    cudaEventRecord(e0);
    k_fma_peak<<<blocks, threads>>>(out, iters, 0.999f, 1e-3f);
    rc = after_launch("k_fma_peak");
    cudaEventRecord(e1);
    cudaError_t e = cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    if (rc) return rc;
    if (e != cudaSuccess) return cuda_check(e, "k_fma_peak sync");
    const double flops = 2.0 * 8 * 16 * (double)iters * blocks * threads;
    *tflops = flops / (ms * 1e-3) / 1e12;
    return QBOLD_OK;
}
