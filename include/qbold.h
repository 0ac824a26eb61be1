/*
 * qbold.h -- C ABI of libqbold.so: the B200-native (sm_100a) implementation of the
 * qBOLD-VI hot path (voxel-wise ASE qBOLD forward signal model fused with the
 * amortized-VI likelihood).
 *
 * The reference (wearepal/qBOLD-VI) has no FFI: the path sits behind Python/Keras
 * objects.  Each entry point below names the reference interface it replaces
 * (file:line in the reference checkout).  The Python host mirror in
 * qbold_vi_b200/ binds these with ctypes and re-exposes the reference's own names
 * (SignalGenerationLayer, create_synthetic_dataset, ReparamTrickLayer,
 * EncoderTrainer.*); see INTEGRATION.md.
 *
 * Conventions
 *   - all array pointers are DEVICE pointers owned by the caller unless the function
 *     name ends in _host; float32, C-contiguous, last axis fastest;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *     calls are asynchronous with respect to the host;
 *   - no hidden device allocation except the *_host helpers (cached staging buffers);
 *   - return value: 0 = ok, negative = QBOLD_E*; qbold_last_error() gives the text
 *     (thread-local);
 *   - QboldParams is an immutable POD; the functions are stateless and may be called
 *     concurrently from one host thread per GPU.
 */
#ifndef QBOLD_H_
#define QBOLD_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QBOLD_MAX_TAU 32
#define QBOLD_NQ 129              /* 2**7 + 1 Simpson nodes, signals.py:168 */
#define QBOLD_NQ_PAD 132
#define QBOLD_ABI_VERSION 2
#define QBOLD_SCHED_PHASE_LEN 4       /* chunks (warp passes) per phase of the static lane schedule */
#define QBOLD_SCHED_MAX_PHASES 10
#define QBOLD_SCHED_MAX_ENTRIES (QBOLD_SCHED_MAX_PHASES * QBOLD_SCHED_PHASE_LEN * 32)

#define QBOLD_OK 0
#define QBOLD_EINVAL (-1)         /* bad argument (the reference would assert / raise) */
#define QBOLD_ECUDA (-2)          /* CUDA runtime error */
#define QBOLD_EUNSUPPORTED (-3)   /* e.g. noise requested for n_tau not in {11,24}: signals.py:117-121 */

/* Physics constants exactly as SignalGenerationLayer.__init__ parses them from the
 * `config` INI (signals.py:29-46), as doubles (Python floats). */
typedef struct QboldPhysics {
    double gamma, b0, dchi, te, r2t, tr, ti, t1b, hct;
} QboldPhysics;

/* Likelihood options of EncoderTrainer (model.py:54-95). */
typedef struct QboldLikelihood {
    int32_t se_idx;                    /* model.py:95 */
    int32_t multi_image_normalisation; /* model.py:540-545 */
    int32_t predict_log_data;          /* model.py:547-549 */
    int32_t reserved;
    double student_t_df;               /* <50 -> StudentT, else Gaussian (model.py:557-561); <=0 = None */
} QboldLikelihood;

/* Resolved, device-ready parameter block (passed by value to the kernels as a
 * __grid_constant__).  Filled by qbold_params_init(); treat as opaque + immutable. */
typedef struct QboldParams {
    int32_t abi_version;
    int32_t n_tau;
    int32_t n_cols;                    /* distinct non-zero |tau| values */
    int32_t full_model;                /* signals.py:192 */
    int32_t include_blood;             /* signals.py:100 */
    int32_t se_idx;
    int32_t multi_image_normalisation;
    int32_t predict_log_data;
    float student_t_df;                /* <=0 or >=50: Gaussian */
    float student_t_logc;              /* lgamma((df+1)/2)-lgamma(df/2)-0.5*log(df*pi) */
    float dw_k_nohct;                  /* float32((4/3) pi gamma b0 dchi)            signals.py:144 */
    float dw_k;                        /* float32((4/3) pi gamma b0 dchi hct)        signals.py:144 */
    float hct;
    float e_tissue;                    /* exp(-te*r2t), float32 exp                  signals.py:172 */
    float kappa;                       /* m_bld * nb (float32 ops)                   signals.py:105-107 */
    float e_blood;                     /* exp(-r2b*te)                               signals.py:241 */
    float blood_c0;                    /* (4/45) hct (1-hct)                         signals.py:239 */
    float blood_c1;                    /* 4 pi b0 dchi                               signals.py:239 */
    float blood_hg;                    /* 0.5 gamma^2                                signals.py:241 */
    float blood_td2;                   /* td^2, td = 2.6^2/2 ms                      signals.py:236-238 */
    float node0_c;                     /* true Simpson weight * g(u_0) of node 0 (value path) */
    float pad0;
    float tau[QBOLD_MAX_TAU];          /* tf.range(...) values, float32              signals.py:34 */
    float blood_b[QBOLD_MAX_TAU];      /* tau-only bracket of calc_blood             signals.py:242-247 */
    float abs_tau[QBOLD_MAX_TAU];      /* distinct non-zero |tau| (bitwise dedup)    */
    int32_t col_of_tau[QBOLD_MAX_TAU]; /* tau index -> column, -1 for tau == 0       */
    float norm_snr[QBOLD_MAX_TAU];     /* signals.py:119-121; all zero if undefined  */
    float qu[QBOLD_NQ_PAD];            /* quadrature nodes u_k (tf.linspace)         signals.py:166-168 */
    float qc[QBOLD_NQ_PAD];            /* value weights  W_k g_k, qc[0]=0 (node 0 dead in FP32), qc[128]=0 */
    float qd[QBOLD_NQ_PAD];            /* derivative weights W_k g_k u_k (node 0 live) */
    /* Static lane schedule of the quadrature (n_cols <= 8; sched_phases == 0 -> column-major path).
     * All (column j, node k>=1) pairs are dealt to the 32 lanes in passes ("chunks") ordered by
     * m = (|tau_j|/tau_ref) * u_k, so that within one pass every lane evaluates the Bessel pair at a
     * similar argument x = A*m (A = 1.5*tau_ref*dw) and the small/mid/large branch is warp-uniform.
     * A lane keeps one column for a whole phase (QBOLD_SCHED_PHASE_LEN passes); see DESIGN.md 3. */
    int32_t sched_phases;
    float tau_ref;                                   /* max |tau| */
    float sched_ph_min[16];                          /* min / max m over the live entries of a phase */
    float sched_ph_max[16];
    float sched_m[QBOLD_SCHED_MAX_ENTRIES];          /* entry [phase][pass][lane]: m */
    float sched_w[QBOLD_SCHED_MAX_ENTRIES];          /*                            Simpson weight c_k (0 = idle) */
    uint8_t sched_col[QBOLD_SCHED_MAX_PHASES * 32];  /* [phase][lane]: column (bits 0-2) | 0x80 = first visit */
} QboldParams;

int qbold_abi_version(void);
/* sizeof(QboldParams) as compiled into the library: bindings verify their struct layout against it. */
int qbold_params_sizeof(void);
const char* qbold_last_error(void);

/* Replaces SignalGenerationLayer.__init__ (signals.py:18-53): resolve the physics,
 * the tau grid (explicit float32 values) and the flags. */
int qbold_params_init(QboldParams* out, const QboldPhysics* phys, const float* taus, int32_t n_tau,
                      int32_t full_model, int32_t include_blood);
/* Sets the EncoderTrainer likelihood options used by qbold_elbo_fused (model.py:54-95). */
int qbold_params_set_likelihood(QboldParams* p, const QboldLikelihood* lik);

/* Replaces SignalGenerationLayer.call with noise off (signals.py:55-114,137-140).
 * oef_dbv [n,width], width 2 (OEF,DBV) or 3 (+Hct, variable_hct); signal [n,n_tau].
 * Domain: OEF >= 0 (every caller of the reference produces OEF in [0.04, 0.84]).  A negative OEF is outside what the
 * scheduled quadrature handles (the build option QB_SIGNED_OEF lifts this): results for such rows
 * are not meaningful and may be non-finite.  NaN / Inf rows give NaN / Inf in that row only. */
int qbold_forward(const QboldParams* p, const float* oef_dbv, int32_t width, int64_t n,
                  float* signal, void* stream);

/* Forward + the vector-Jacobian product TensorFlow autodiff gives through
 * SignalGenerationLayer.call (bessel_j0' = -bessel_j1): g_oef_dbv[n,2] =
 * d(sum(signal*g_signal))/d(oef,dbv).  g_signal NULL means all ones (the
 * signals.py:307-314 demo).  signal may be NULL. */
int qbold_forward_backward(const QboldParams* p, const float* oef_dbv, const float* g_signal,
                           int64_t n, float* signal, float* g_oef_dbv, void* stream);

/* variable_hct=True (signals.py:64-70): rows are (OEF, DBV, Hct); g_oef_dbv_hct[n,3] also carries
 * d/dHct (through dw = (4/3) pi gamma B0 dchi Hct OEF, signals.py:142-144, and the blood term's
 * G0 ~ Hct (1 - Hct), signals.py:239). */
int qbold_forward_backward_hct(const QboldParams* p, const float* oef_dbv_hct, const float* g_signal,
                               int64_t n, float* signal, float* g_oef_dbv_hct, void* stream);

/* Misalignment augmentation of SignalGenerationLayer.call (signals.py:80-96), applied to `signal`
 * [n,n_tau] = the noise-free forward model of oef_dbv [n,width]: voxels with sel_u01 < prob get, for
 * the images after from_index (in [4, n_tau-1)), the signal of OEF + 0.15 eps[.,0] clipped to
 * [0.05,0.8] and DBV + 0.05 eps[.,1] clipped to [0.002,0.3].  sel_u01 [n], from_index [n] (int32),
 * eps [n,2]: the reference's draws (all three or none); NULL = Philox (seed, offset + voxel,
 * stream 0x20000).  QBOLD_EUNSUPPORTED when n_tau <= 5 (the reference's randint range is empty). */
int qbold_misalign(const QboldParams* p, const float* oef_dbv, int32_t width, int64_t n, float prob,
                   const float* sel_u01, const int32_t* from_index, const float* eps, uint64_t seed,
                   uint64_t offset, float* signal, void* stream);

/* Same call with HOST buffers: chunked, double-buffered H2D -> kernel -> D2H on
 * internal streams of the current device; returns after the results are in host memory. */
int qbold_forward_backward_host(const QboldParams* p, const float* h_oef_dbv, const float* h_g_signal,
                                int64_t n, float* h_signal, float* h_g_oef_dbv);

/* The host <-> device ceiling of this box for the copy pattern of qbold_forward_backward_host (same chunk size,
 * same number of streams, H2D and D2H concurrently, no kernel).  h_src, h_dst: host buffers of `bytes` each
 * (pinned); gbps[0] = H2D GB/s, gbps[1] = D2H GB/s while both directions run (best of `reps`). */
int qbold_host_copy_ceiling(const void* h_src, void* h_dst, int64_t bytes, int32_t reps, double* gbps);

/* Replaces ReparamTrickLayer.call (model.py:21-50), use_mvg branch: q [n,5] raw encoder
 * outputs; eps [n,2] explicit N(0,1) draws, or NULL -> Philox4x32-10 keyed (seed, voxel). */
int qbold_reparam_sample(const float* q, const float* eps, uint64_t seed, uint64_t offset,
                         int64_t n, float* oef_dbv, void* stream);

/* Column means of a [n,n_tau] signal block (the batch statistic of signals.py:126);
 * mean is a device array [n_tau]; scratch is a device array of >= 2*n_tau doubles. */
int qbold_column_mean(const float* signal, int64_t n, int32_t n_tau, float* mean, double* scratch,
                      void* stream);

/* Noise model (signals.py:116-128) in place: snr ~ U(50,120) * norm_snr, std = mean/snr,
 * signal += N(0,1)*std.  snr_u01 [n] and eps [n,n_tau] are explicit draws, or both NULL ->
 * Philox keyed by (seed, offset+voxel). */
int qbold_add_noise(const QboldParams* p, float* signal, int64_t n, const float* mean,
                    const float* snr_u01, const float* eps, uint64_t seed, uint64_t offset,
                    void* stream);

/* The noise loop of create_synthetic_dataset (signals.py:282-285) for all chunks at once: signal [n_chunks*chunk_rows,
 * n_tau] in place, every chunk uses ITS OWN column means, row v draws with counter offset + v.  scratch: 32*n_chunks
 * doubles of device memory.  Same draws and arithmetic as qbold_column_mean + qbold_add_noise per chunk. */
int qbold_add_noise_chunked(const QboldParams* p, float* signal, int64_t chunk_rows, int32_t n_chunks,
                            const float* snr_u01, const float* eps, uint64_t seed, uint64_t offset, double* scratch,
                            void* stream);

/* Replaces the body of create_synthetic_dataset (signals.py:270-299) for rows
 * [first, first+count) of the shuffled OEF x DBV meshgrid: labels y3 = (OEF, DBV, R2')
 * and the clean signal x.  perm (int64 [n_oef*n_dbv]) is an explicit shuffle, or NULL ->
 * keyed Feistel bijection of the index space (no permutation array in HBM). */
int qbold_generate(const QboldParams* p, const float* oefs, int64_t n_oef, const float* dbvs,
                   int64_t n_dbv, const int64_t* perm, uint64_t seed, int64_t first, int64_t count,
                   float* x, float* y3, void* stream);

/* Fused training step of the VI likelihood: replaces ReparamTrickLayer (model.py:21-50)
 * -> SignalGenerationLayer (signals.py:55-114) -> fine_tune_loss_fn (model.py:527-568)
 * + kl_weight * kl_loss (model.py:654-665, 70-sample MC estimator :592-610) and the
 * backward pass TensorFlow autodiff would run, in one kernel.
 *   q, prior [n,5]; sigma, y [n,n_tau]; mask [n];
 *   eps [n,2] / eps_kl [n,kl_samples,2] explicit draws or NULL -> Philox (seed, offset+voxel);
 *   inv_mask_sum = 1/sum(mask) over the GLOBAL batch (all ranks);
 *   outputs: grad_q [n,5], grad_sigma [n,n_tau] (d(nll + kl_weight*kl)/d.), optional
 *   nll_map/kl_map [n], sums[4] (device doubles: sum nll*mask, sum kl, sum mask, #non-finite). */
int qbold_elbo_fused(const QboldParams* p, const float* q, const float* sigma, const float* y,
                     const float* mask, const float* prior, const float* eps, const float* eps_kl,
                     uint64_t seed, uint64_t offset, int32_t kl_samples, float inv_mask_sum,
                     float kl_weight, int64_t n, float* grad_q, float* grad_sigma, float* nll_map,
                     float* kl_map, double* sums, void* stream);

/* Same kernel with 1/sum(mask) read from DEVICE memory (one float): the data-parallel trainer all-reduces the mask
 * count on the device and never brings it to the host, so a training step has no host synchronisation. */
int qbold_elbo_fused_dev(const QboldParams* p, const float* q, const float* sigma, const float* y,
                         const float* mask, const float* prior, const float* eps, const float* eps_kl,
                         uint64_t seed, uint64_t offset, int32_t kl_samples, const float* inv_mask_sum_dev,
                         float kl_weight, int64_t n, float* grad_q, float* grad_sigma, float* nll_map,
                         float* kl_map, double* sums, void* stream);

/* Same kernel for a CAPTURED training step (CUDA graph): every per-step scalar comes from device memory, so the
 * launch replays with fresh draws.  seed_dev: one uint64, the Philox key (the trainer's per-call seed,
 * qbold_elbo_fused's `seed`), advanced on the device between replays; draws are always in-kernel (no eps / eps_kl). */
int qbold_elbo_fused_graph(const QboldParams* p, const float* q, const float* sigma, const float* y,
                           const float* mask, const float* prior, const uint64_t* seed_dev, uint64_t offset,
                           int32_t kl_samples, const float* inv_mask_sum_dev, float kl_weight, int64_t n,
                           float* grad_q, float* grad_sigma, float* nll_map, float* kl_map, double* sums,
                           void* stream);

/* fine_tune_loss_fn alone (model.py:527-568) for predictions already in HBM: nll_map[n] = mask * sum_tau NLL,
 * and (optional) d_pred / d_sigma [n,n_tau] = d nll_map[v] / d pred[v,:], d sigma[v,:].  mask may be NULL. */
int qbold_nll(const QboldParams* p, const float* y, const float* pred, const float* sigma, const float* mask,
              int64_t n, float* nll_map, float* d_pred, float* d_sigma, void* stream);

/* kl_loss alone (model.py:654-665 -> mvg_kl_samples :592-610): per-voxel KL(q || prior) map
 * (zero where mask <= 0; mask may be NULL) and, optionally, grad_q[n,5] = d kl_map[v] / d q[v,:]
 * (path-derivative estimator: stop_gradient on q inside log q, model.py:596).
 * n_samples = 70 is the reference default; n_samples = 0 selects the closed-form KL. */
int qbold_kl(const float* q, const float* prior, const float* mask, const float* eps_kl, uint64_t seed,
             uint64_t offset, int32_t n_samples, int64_t n, float* kl_map, float* grad_q, void* stream);

/* Replaces calculate_means(include_r2p=True, return_stds=True) (model.py:326-343):
 * n_samples reparameterised draws per voxel -> mean3/var3 [n,3] of (OEF, DBV, R2').
 * eps [n,n_samples,2] or NULL -> Philox. */
int qbold_posterior_stats(const QboldParams* p, const float* q, const float* eps, uint64_t seed,
                          uint64_t offset, int32_t n_samples, int64_t n, float* mean3, float* var3,
                          void* stream);

/* Posterior-predictive likelihood map of save_predictions (model.py:808-817): mean over n_samples
 * reparameterised draws of fine_tune_loss_fn(..., return_mean=False), forward-only.  eps [n,n_samples,2] or
 * NULL -> Philox; mask may be NULL (all ones); nll_map [n]. */
int qbold_nll_map(const QboldParams* p, const float* q, const float* sigma, const float* y, const float* mask,
                  const float* eps, uint64_t seed, uint64_t offset, int32_t n_samples, int64_t n, float* nll_map,
                  void* stream);

/* ---- losses either side of the fused path ---------------------------------------- */

/* smoothness_loss (model.py:726-754): total variation of the forward-transformed, range-rescaled means over
 * x / y neighbours that are both inside the mask.  q [n_vol,nx,ny,nz,n_ch] (n_ch = 5 mvg / 4 diagonal; channels 0
 * and 2 are the means), mask [n_vol,nx,ny,nz].  *tv_sum += sum |d| (double, caller zeroes; may be NULL);
 * grad_q (same shape as q, may be NULL) = scale * d(sum |d|)/dq, so scale = weight / sum(mask). */
int qbold_smoothness(const float* q, int32_t n_ch, const float* mask, int64_t n_vol, int32_t nx, int32_t ny,
                     int32_t nz, float scale, double* tv_sum, float* grad_q, void* stream);
/* Same kernel with `scale` read from device memory (see qbold_elbo_fused_dev). */
int qbold_smoothness_dev(const float* q, int32_t n_ch, const float* mask, int64_t n_vol, int32_t nx, int32_t ny,
                         int32_t nz, const float* scale_dev, double* tv_sum, float* grad_q, void* stream);

/* synthetic_data_loss (model.py:449-514) per label row: logit-MVN NLL of (OEF, DBV) under the predicted
 * distribution (use_mvg: logit_gaussian_mvg_log_prob :376-400, pred [n,5]; else logit_gaussian_log_prob :406-421,
 * pred [n,4]) minus, when inv_gamma_alpha*inv_gamma_beta > 0, the InverseGamma log-prior of the predicted
 * variances (:495-507).  labels [n,label_stride] (OEF, DBV first).  Outputs (each may be NULL): nll_rows [n],
 * grad_pred = grad_scale * d nll_rows[v] / d pred[v,:], *loss_sum += sum of rows (double, caller zeroes). */
int qbold_synth_nll(const float* labels, int32_t label_stride, const float* pred, int32_t use_mvg,
                    double inv_gamma_alpha, double inv_gamma_beta, int64_t n, float grad_scale, float* nll_rows,
                    float* grad_pred, double* loss_sum, void* stream);

/* synthetic_data_loss with infer_inv_gamma=True (model.py:454-455,493-496): the InverseGamma parameters are
 * LEARNED.  inv_gamma_params (DEVICE pointer, 4 floats: alpha_oef, beta_oef, alpha_dbv, beta_dbv -- the exp() of the
 * encoder's hyper-prior variables, model.py:201-205).  pred rows have stride pred_stride (the reference layout is
 * [q | 4 hyper-prior channels]); grad_pred is [n, 5|4] dense.  ig_sums (double[4], caller zeroes, may be NULL) +=
 * (sum log v_oef, sum 1/v_oef, sum log v_dbv, sum 1/v_dbv), from which d loss / d(alpha, beta) follow in closed
 * form: d/dalpha = -(n (log beta - digamma(alpha)) - sum log v), d/dbeta = -(n alpha / beta - sum 1/v). */
int qbold_synth_nll_inferred(const float* labels, int32_t label_stride, const float* pred, int32_t pred_stride,
                             int32_t use_mvg, const float* inv_gamma_params, int64_t n, float grad_scale,
                             float* nll_rows, float* grad_pred, double* loss_sum, double* ig_sums, void* stream);

/* Mixture-of-Gaussians population prior of kl_loss (model.py:666-684; use_mvg=False, mog_components = M > 1):
 * pred [n, 4 (M+1)] = q then M components (mean, raw std of OEF and DBV each); single-sample estimate with the
 * draws eps [n,2] (NULL = Philox (seed, offset + voxel, stream 0)).  kl_map [n] (0 where mask <= 0; mask may be
 * NULL), grad_pred [n, 4 (M+1)] (may be NULL) = d kl_map[v] / d pred[v,:]. */
int qbold_mog_kl(const float* pred, int32_t n_components, const float* mask, const float* eps, uint64_t seed,
                 uint64_t offset, int64_t n, float* kl_map, float* grad_pred, void* stream);

/* KL of the diagonal (use_mvg=False) branch of kl_loss (model.py:685-708): tfp LogitNormal.kl_divergence for OEF
 * plus DBV, zero where mask <= 0 (mask may be NULL).  pred / prior rows = [mean_o, raw_std_o, mean_d, raw_std_d]
 * at the given row strides (the population-prior layout keeps both in one 8-channel tensor, :687-689).
 * grad_pred / grad_prior (may be NULL) = d kl_map[v] / d row, written at their own row strides. */
int qbold_diag_kl(const float* pred, int32_t pred_stride, const float* prior, int32_t prior_stride,
                  const float* mask, int64_t n, float* kl_map, float* grad_pred, int32_t grad_pred_stride,
                  float* grad_prior, int32_t grad_prior_stride, void* stream);

/* ---- stream 1 of the amortization network on the tensor cores ------------------------------ */

/* The voxel-wise branch of create_encoder (model.py:122-223): normalise_data (:97-113) -> Dense(n_in -> H) -> ReLU ->
 * n_mid x [Dense(H -> H) -> ReLU] -> Dense(H -> n_out), one tcgen05 (kind::tf32, fp32 accumulate in TMEM) kernel,
 * inference only.  Limits: n_in <= 32, H <= 64, 1 <= n_mid <= 6, n_out <= 16, ReLU.
 * qbold_encoder_mlp_pack turns torch-layout device weights ([out, in] row-major, biases [out]) into the
 * shared-memory image (`blob`, qbold_encoder_mlp_blob_floats(n_mid) floats, device) the kernel copies verbatim;
 * w_mid / b_mid are HOST arrays of n_mid device pointers.  qbold_encoder_mlp_forward: data [n, n_in] raw images ->
 * q [n, n_out]; *status (device int, may be NULL) becomes non-zero if a tensor-core completion timed out. */
int qbold_encoder_mlp_blob_floats(int32_t n_mid);
int qbold_encoder_mlp_pack(const float* w_in, const float* b_in, const float* const* w_mid, const float* const* b_mid,
                           const float* w_out, const float* b_out, int32_t n_in, int32_t n_hidden, int32_t n_mid,
                           int32_t n_out, float* blob, void* stream);
int qbold_encoder_mlp_forward(const float* data, const float* blob, int32_t n_in, int32_t n_mid, int32_t n_out,
                              int32_t se_idx, int32_t multi_image_normalisation, int64_t n, float* q, int32_t* status,
                              void* stream);

/* Weight / bias gradient of a per-voxel Dense layer (the 1x1x1 convolutions of create_encoder, model.py:122-223):
 * dw[n_out,n_in] = sum_v g'[v,:]^T x[v,:], db[n_out] = sum_v g'[v,:] (db may be NULL), g' = g * [relu_mask > 0];
 * TF32 mma, fp32 accumulate,
 * deterministic two-stage reduction.  n_out <= 64, n_in <= 63.  workspace: qbold_dense_wgrad_workspace_floats()
 * floats of device scratch.  accumulate != 0 adds into dw / db. */
int64_t qbold_dense_wgrad_workspace_floats(void);
int qbold_dense_wgrad(const float* g, const float* relu_mask, int32_t n_out, const float* x, int32_t n_in, int64_t n,
                      float* dw, float* db, int32_t accumulate, float* workspace, void* stream);

/* Skinny Dense layers (the encoder's heads, n_out <= 16: 60 -> 5 posterior parameters, 60 -> 11 sigmas) as streaming
 * FP32 passes: y[n,n_out] = x[n,n_in] w^T + bias and dx[n,n_in] = g[n,n_out] w, w [n_out,n_in] row-major; n_in a multiple
 * of 4 up to 64; x / w / y / dx 16-byte aligned. */
int qbold_dense_small_forward(const float* x, const float* w, const float* bias, int32_t n_in, int32_t n_out, int64_t n,
                              float* y, void* stream);
int qbold_dense_small_dgrad(const float* g, const float* w, int32_t n_in, int32_t n_out, int64_t n, float* dx,
                            void* stream);

/* dx[n,n_in] = [relu_mask > 0] * (g[n,n_out] w): the skinny input gradient with the ReLU' of the activation the head
 * read applied to the result (relu_mask [n,n_in], may be NULL); warp-cooperative coalesced stores.  n_out <= 16,
 * n_in <= 64. */
int qbold_dense_small_dgrad_masked(const float* g, const float* w, const float* relu_mask, int32_t n_in, int32_t n_out,
                                   int64_t n, float* dx, void* stream);

/* One Dense layer on the tensor cores (tcgen05 kind::tf32, fp32 accumulate) for the encoder's training passes:
 * y[n,n_out] = act((x[n,n_in] * [relu_mask > 0]) B^T + bias).  qbold_dense_tc_pack builds the operand image
 * (qbold_dense_tc_packed_floats() floats, device) from a row-major matrix: transpose = 0: B = w [rows, cols] with
 * bias [rows] (forward); transpose = 1: B = w^T where w is stored [cols, rows] (input gradient, bias NULL).
 * relu_mask (may be NULL) has the shape of x: the layer's ReLU output, fusing ReLU' into the input-gradient pass.
 * n_in, n_out: multiples of 4 in [4, 64]; x, y, relu_mask 16-byte aligned.  relu_mask of qbold_dense_wgrad (may be
 * NULL, shape of g) does the same for the weight gradient. */
int qbold_dense_tc_packed_floats(void);
int qbold_dense_tc_pack(const float* w, const float* bias, int32_t rows, int32_t cols, int32_t transpose, float* packed,
                        void* stream);
int qbold_dense_tc(const float* x, const float* relu_mask, const float* packed, int32_t n_in, int32_t n_out, int32_t relu,
                   int64_t n, float* y, int32_t* status, void* stream);

/* Gated residual mix of the encoder blocks (model.py:160-172): out = skip*(1-g) + r*g, g = sigmoid(z + offset);
 * skip, r, out [n, channels]; z [n, z_channels] with z_channels = channels (channel-wise gating) or 1.
 * Backward: d_skip, d_r [n, channels], d_z [n, z_channels] from the upstream gradient go. */
int qbold_gate_mix_forward(const float* skip, const float* r, const float* z, float offset, int64_t n,
                           int32_t channels, int32_t z_channels, float* out, void* stream);
int qbold_gate_mix_backward(const float* go, const float* skip, const float* r, const float* z, float offset, int64_t n,
                            int32_t channels, int32_t z_channels, float* d_skip, float* d_r, float* d_z, void* stream);

/* Fused elementwise steps of the encoder block's training path (create_block, model.py:142-174; see
 * csrc/encoder_block.cu).  channels: a multiple of 4 up to 64, channel-wise gating; operands 16-byte aligned.
 *   forward : out = skip (1-g) + (r0 + r_bias) g, g = sigmoid(z + offset); r_bias [channels] may be NULL (r0 is
 *             the second 3x3x1 convolution without its bias); out_relu (may be NULL) = relu(out), the next block's
 *             convolution input (model.py:150).
 *   backward: d_r, d_z as qbold_gate_mix_backward; d_skip is multiplied by [skip > 0] when skip_is_relu != 0 (the
 *             skip branch ends in a ReLU). */
int qbold_block_mix_forward(const float* skip, const float* r0, const float* r_bias, const float* z, float offset,
                            int64_t n, int32_t channels, float* out, float* out_relu, void* stream);
int qbold_block_mix_backward(const float* go, const float* skip, const float* r0, const float* r_bias, const float* z,
                             float offset, int64_t n, int32_t channels, int32_t skip_is_relu, float* d_skip, float* d_r,
                             float* d_z, void* stream);
/* The same with a second gradient of the skip activation (skip_addend [n, channels], may be NULL) added BEFORE the ReLU'
 * mask: d_skip = [skip > 0] * (go (1 - g) + skip_addend) -- block 0 of the encoder, whose stream-1 output IS the skip. */
int qbold_block_mix_backward_add(const float* go, const float* skip, const float* r0, const float* r_bias, const float* z,
                                 float offset, int64_t n, int32_t channels, int32_t skip_is_relu, const float* skip_addend,
                                 float* d_skip, float* d_r, float* d_z, void* stream);
/* ReLU backward fused with the bias gradient: out[n,channels] = g * [y > 0] (+ addend, may be NULL) (y NULL: no mask,
 * out unused) and colsum[channels] (+)= column sums of that (colsum may be NULL; deterministic two-stage reduction
 * through `workspace`, qbold_colsum_workspace_floats() floats of device scratch). */
int64_t qbold_colsum_workspace_floats(void);
int qbold_relu_bwd_colsum(const float* g, const float* y, const float* addend, int64_t n, int32_t channels, float* out,
                          float* colsum, int32_t accumulate, float* workspace, void* stream);

/* normalise_data (model.py:97-113) fused with the training path's layout change: raw images data [b, nx, ny, nz, n_tau]
 * -> out [b, nz, nx, ny, tp] = log(clip(d, 1e-2, 1e8) / reference), tp = n_tau rounded up to a multiple of 4, padding
 * columns zero.  reference = the clipped tau = 0 image (se_idx) or the mean of the three clipped images around it. */
int qbold_normalise_zouter(const float* data, int64_t b, int32_t nx, int32_t ny, int32_t nz, int32_t n_tau, int32_t se_idx,
                           int32_t multi_image_normalisation, float* out, void* stream);

/* One Dense layer as a persistent TMA -> tcgen05 (kind::tf32, fp32 accumulate in TMEM) -> TMA pipeline, for the
 * encoder's training passes (reference create_layer, model.py:115-120):
 *     y[n, n_out] = act(x[n, n_in] B + bias) (+ addend[n, n_out])
 * transpose = 0: B = w^T with w [n_out, n_in] (forward); transpose = 1: B = w with w [n_in, n_out] (input gradient: the
 * stored weight is read MN-major, no transposed copy).  bias, addend may be NULL; addend may alias y (beta = 1
 * accumulation in place).  n_in, n_out: multiples of 4 in [4, 64]; all matrices dense row-major, 16-byte aligned.
 * status (may be NULL) is set non-zero if a completion barrier timed out. */
int qbold_dense_tma(const float* x, const float* w, const float* bias, const float* addend, int32_t n_in, int32_t n_out,
                    int32_t transpose, int32_t relu, int64_t n, float* y, int32_t* status, void* stream);

/* Weight / bias gradient of a Dense layer on the same machinery: dw[n_out, n_in] (+)= g^T x, db[n_out] (+)= column sums
 * of g (db may be NULL); g [n, n_out], x [n, n_in] read MN-major straight from memory by TMA (the contraction runs over
 * the rows), one TMEM accumulator per SM, per-SM partials in `workspace` (qbold_dense_wgrad_tma_workspace_floats()
 * floats) summed in a fixed order.  n_in, n_out: multiples of 4 in [4, 64]. */
int64_t qbold_dense_wgrad_tma_workspace_floats(void);
int qbold_dense_wgrad_tma(const float* g, int32_t n_out, const float* x, int32_t n_in, int64_t n, float* dw, float* db,
                          int32_t accumulate, float* workspace, int32_t* status, void* stream);

/* Weight gradient of the encoder's 3x3x1 convolutions (create_block, model.py:152,156; padding 'same') on tcgen05
 * tensor cores (kind::tf32, fp32 accumulate in TMEM): dw[cg, cx, 3, 3] (+)= sum_v g[v, :]^T x[v + (kx-1, ky-1), :]
 * for activations laid out [n_images, nx, ny, channels] (the encoder's z-outer layout: n_images = B * Z).
 * cg, cx: multiples of 4 up to 64; g, x 16-byte aligned.  workspace: qbold_conv_wgrad_workspace_floats() floats of
 * device scratch (per-SM partials, summed in a fixed order).  status (may be NULL) is set non-zero if a tensor-core
 * completion barrier timed out. */
int64_t qbold_conv_wgrad_workspace_floats(void);
int qbold_conv_wgrad(const float* g, int32_t cg, const float* x, int32_t cx, int64_t n_images, int32_t nx, int32_t ny,
                     float* dw, int32_t accumulate, float* workspace, int32_t* status, void* stream);

/* FP32 FMA micro-benchmark (roofline denominator measured in the same run): launches
 * `iters` dependent-chain FFMA sweeps, returns achieved TFLOP/s through *tflops. */
int qbold_fma_peak(int32_t iters, double* tflops);

/* Counters: number of kernels this library launched since load (bench's gpu_launches). */
int64_t qbold_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* QBOLD_H_ */
