"""NumPy restatement of the qBOLD-VI hot path.  TEST INFRASTRUCTURE ONLY.

Every function cites the reference lines it follows (paths are relative to the
reference checkout, e.g. ``signals.py:159-172``).  ``dtype=np.float32`` rounds
every elementary operation to float32 in the reference's own order of operations
(TensorFlow semantics: Python scalars are converted to the tensor dtype before the
op); ``dtype=np.float64`` evaluates the same formulas in double precision with
scipy Bessel functions and reproduces the single FP32 artefact that is part of
the reference's results: quadrature node 0 contributes exactly 0 to the value but
is live in the derivative.

All randomness enters as explicit arrays (``eps``) -- TensorFlow's Philox streams
cannot be reproduced offline.

PARITY PINNING: see oracle/__init__.py ("parity unpinned" w.r.t. real TensorFlow;
pinned against the reference source run over oracle/tf_shim and App. B KATs).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import scipy.special as _sp

from . import cephes_f32

F32 = np.float32
F64 = np.float64

# model.py:88-91
OEF_RANGE, MIN_OEF, DBV_RANGE, MIN_DBV = 0.8, 0.04, 0.2, 0.001
NB = 0.775                     # signals.py:102
N_QUAD = 2 ** 7 + 1            # signals.py:168
NORM_SNR_11 = np.array([0.985, 1.00, 1.01, 1., 0.97, 0.95, 0.93, 0.90, 0.86, 0.83, 0.79],
                       dtype=np.float32)                          # signals.py:119


# --------------------------------------------------------------------------
# parameters (signals.py:29-46)
# --------------------------------------------------------------------------
@dataclass
class Physics:
    gamma: float
    b0: float
    dchi: float
    te: float
    r2t: float
    tr: float
    ti: float
    t1b: float
    hct: float
    taus: np.ndarray                       # float32 [n_tau]
    simulate_noise: bool = False
    extra: dict = field(default_factory=dict)

    @property
    def n_tau(self):
        return int(self.taus.shape[0])


def make_taus(tau_start, tau_end, tau_step):
    """tf.range(start, limit, delta, dtype=float32) (signals.py:34-35).

    size = ceil(|limit-start|/|delta|) and element i = start + i*delta, all in
    float32 (TensorFlow's RangeOp; older releases accumulate ``val += delta`` -- a
    <=1 ulp difference, SURVEY.md A.6 item 13 -- hence tau is always an explicit
    input of the kernels)."""
    s, e, d = F32(tau_start), F32(tau_end), F32(tau_step)
    n = int(math.ceil(abs(float((e - s) / d))))
    return (s + np.arange(n, dtype=np.float32) * d).astype(np.float32)


def parse_params(system_parameters, taus=None) -> Physics:
    """signals.py:29-46 -- the mapping holds *strings*; booleans are the literal 'True'."""
    sp = system_parameters
    if taus is None:
        taus = make_taus(float(sp['tau_start']), float(sp['tau_end']), float(sp['tau_step']))
    return Physics(gamma=float(sp['gamma']), b0=float(sp['b0']), dchi=float(sp['dchi']),
                   te=float(sp['te']), r2t=float(sp['r2t']), tr=float(sp['tr']), ti=float(sp['ti']),
                   t1b=float(sp['t1b']), hct=float(sp['hct']),
                   taus=np.asarray(taus, dtype=np.float32),
                   simulate_noise=(str(sp.get('simulate_noise', 'False')) == 'True'))


def default_config():
    """The reference's ``config`` INI ([DEFAULT], config:1-61) as a dict of strings."""
    return {
        'tr': '3.0', 'ti': '1.21', 'te': '0.074', 'tau_start': '-0.016', 'tau_end': '0.065',
        'tau_step': '0.008', 'dchi': '2.64e-7', 'gamma': '2.67513e8', 'b0': '3.0', 't1b': '1.58',
        'r2t': '11.5', 'td': '3.755555555', 'nb': '0.775', 'hct': '0.34', 's0': '100',
        'simulate_noise': 'True', 'tau_weighted': 'True', 'snr': '10',
        'oef_start': '0.05', 'oef_end': '0.8', 'oef_mean': '0.4', 'oef_std': '0.2',
        'dbv_start': '0.003', 'dbv_end': '0.195', 'dbv_mean': '0.025', 'dbv_std': '0.02',
        'sample_size': '2500',
    }


# --------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------
def _c(v, dt):
    """Python scalar -> tensor dtype (TF converts scalars before the op)."""
    return dt(v)


def _exp(x, dt):
    return np.exp(np.asarray(x, dtype=dt)).astype(dt)


def quad_nodes():
    """tf.linspace(1e-5, 1, 129) in float32 (signals.py:166-168): u_0=start,
    u_i = start + delta*i, u_128 = stop, delta = (stop-start)/128."""
    a, b = F32(1e-5), F32(1.0)
    delta = (b - a) / F32(N_QUAD - 1)
    u = a + delta * np.arange(N_QUAD, dtype=np.float32)
    u[0], u[-1] = a, b
    return u.astype(np.float32)


def simpson_weights(u, dt):
    """integral() (signals.py:174-185): sum over 64 panels of (y_a+y_b+4y_m)*(h/3),
    h=(u[2]-u[0])/2 -> per-node weights h/3*[1,4,2,...,4,1]."""
    u = u.astype(dt)
    h = (u[2] - u[0]) / dt(2.0)
    w = np.full(N_QUAD, 2.0, dtype=dt)
    w[1::2] = 4.0
    w[0] = w[-1] = 1.0
    return (w * (h / dt(3.0))).astype(dt), h


def dw_const(ph: Physics, hct=None):
    """calculate_dw_static (signals.py:142-144): a Python-double product, then * tensor."""
    hct = ph.hct if hct is None else hct
    return (4.0 / 3.0) * math.pi * ph.gamma * ph.b0 * ph.dchi * hct


def _j0(x, dt):
    return cephes_f32.j0f(x) if dt is F32 else _sp.j0(x)


def _j1(x, dt):
    return cephes_f32.j1f(x) if dt is F32 else _sp.j1(x)


def _one_minus_j0(arg, dt):
    """1 - bessel_j0(arg) (signals.py:170).  In float64 mode the float32 artefact
    of node 0 is applied by the caller (see _node0_dead)."""
    return (dt(1.0) - _j0(arg, dt)).astype(dt)


def _node0_dead(arg0):
    """True where float32 evaluates 1 - j0f(arg0) to exactly 0: the Cephes tiny
    branch 1 - 0.25*z rounds to 1.0f when 0.25*z <= 2^-25 (ties-to-even)."""
    z = np.asarray(arg0, dtype=np.float64) ** 2
    return 0.25 * z <= 2.0 ** -25


# --------------------------------------------------------------------------
# forward model
# --------------------------------------------------------------------------
def blood_b_of_tau(ph: Physics, dt):
    """The tau-only bracket of calc_blood (signals.py:241-247)."""
    td = ((2.6 ** 2.0) / 2.0) * 1e-3                              # :236-238
    taus = ph.taus.astype(dt)
    t0 = dt(ph.te / td) + np.sqrt(dt(0.25 + (ph.te / td)))
    t0 = t0 + dt(1.5)
    s1 = dt(2.0) * np.sqrt(dt(0.25) + (dt(ph.te) + taus) / dt(td))
    s2 = dt(2.0) * np.sqrt(dt(0.25) + (dt(ph.te) - taus) / dt(td))
    return ((t0 - s1) - s2).astype(dt), td


def m_bld(ph: Physics, dt):
    """signals.py:105 (tf.math.exp on Python floats -> float32 tensors)."""
    e1 = _exp(-(ph.tr - ph.ti) / ph.t1b, dt)
    e2 = _exp(-ph.ti / ph.t1b, dt)
    return dt(dt(1.0) - (dt(2.0) - e1) * e2)


def calc_tissue(ph: Physics, oef, dbv, hct=None, full_model=True, dtype=F32, chunk=4096,
                want_grad=False, g_out=None, want_hct=False):
    """calc_tissue (signals.py:152-209).

    oef, dbv: [N] arrays.  Returns S_t [N, n_tau]; with ``want_grad`` also the
    TF-autodiff-consistent (g_oef, g_dbv) for the upstream gradient ``g_out``
    [N, n_tau] (bessel_j0' = -bessel_j1, node 0 live)."""
    dt = dtype
    oef = np.asarray(oef, dtype=dt).reshape(-1)
    dbv = np.asarray(dbv, dtype=dt).reshape(-1)
    n = oef.shape[0]
    taus = ph.taus.astype(dt)
    nt = taus.shape[0]
    K = dw_const(ph, hct)
    if np.ndim(K) > 0:                                            # variable hct: per-voxel constant
        K = np.asarray(K, dtype=dt).reshape(-1)
    else:
        K = dt(K)
    E = _exp(-ph.te * ph.r2t, dt)                                 # :172 / :204
    out = np.empty((n, nt), dtype=dt)
    g_oef = np.zeros(n, dtype=dt) if want_grad else None
    g_dbv = np.zeros(n, dtype=dt) if want_grad else None
    g_hct = np.zeros(n, dtype=dt) if want_grad else None          # d/d hct through dw = K0 * hct * oef (:142-144)
    K0 = dt(dw_const(ph, 1.0))

    if not full_model:
        # log-linear branch, signals.py:194-207
        dw = (K * oef).astype(dt)
        with np.errstate(divide="ignore"):
            tc = dt(1.0) / dw
        r2p = dw * dbv
        under = (np.abs(taus)[None, :] < tc[:, None]).astype(dt)
        over = dt(1.0) - under
        rt = r2p[:, None] * taus[None, :]
        s = E * np.exp(-(dt(0.3) * rt ** 2) / dbv[:, None])
        s2 = E * np.exp(dbv[:, None] - rt)
        out[:] = s * under + s2 * over
        if want_grad:
            g = np.asarray(g_out, dtype=dt)
            gs, gs2 = g * under, g * over
            # s = E*exp(-0.3*rt^2/dbv)
            g_e1 = gs * s                                         # d/d(exponent)
            g_rt = g_e1 * (-(dt(0.3) * dt(2.0) * rt) / dbv[:, None])
            g_dbv_a = (g_e1 * (dt(0.3) * rt ** 2) / (dbv[:, None] ** 2)).sum(-1)
            # s2 = E*exp(dbv - rt)
            g_e2 = gs2 * s2
            g_dbv_b = g_e2.sum(-1)
            g_rt = g_rt - g_e2
            g_r2p = (g_rt * taus[None, :]).sum(-1)
            g_dbv[:] = g_dbv_a + g_dbv_b + g_r2p * dw
            g_oef[:] = g_r2p * dbv * K
            g_hct[:] = g_r2p * dbv * (K0 * oef)
        if want_grad:
            return (out, g_oef, g_dbv, g_hct) if want_hct else (out, g_oef, g_dbv)
        return out

    u32 = quad_nodes()
    u = u32.astype(dt)
    W, _h = simpson_weights(u32, dt)
    A = (dt(2.0) + u) * np.sqrt(dt(1.0) - u)                      # :169
    D = dt(3.0) * np.square(u)                                    # :171
    for s0 in range(0, n, chunk):
        sl = slice(s0, min(n, s0 + chunk))
        Ks = K[sl] if np.ndim(K) > 0 else K
        dw = (Ks * oef[sl]).astype(dt)                            # :187
        t = taus[None, :] * dw[:, None]                           # taus * dw_i       [m, nt]
        arg = (dt(1.5) * t)[:, :, None] * u[None, None, :]        # 1.5*(..)*u        [m, nt, 129]
        omj = _one_minus_j0(arg, dt)
        if dt is F64:
            dead = _node0_dead(arg[:, :, 0])
            omj[:, :, 0] = np.where(dead, 0.0, omj[:, :, 0])
        y = (A[None, None, :] * omj) / D[None, None, :]           # :169-171
        # integral(): (y_a + y_b + 4 y_m) * (h/3), reduce_sum over the 64 panels
        h3 = _h / dt(3.0)
        panels = ((y[:, :, 0:-2:2] + y[:, :, 2::2]) + dt(4.0) * y[:, :, 1:-1:2]) * h3
        I = panels.sum(-1, dtype=dt)
        st = np.exp(-dbv[sl][:, None] * I).astype(dt) * E         # :169,172
        out[sl] = st
        if want_grad:
            g = np.asarray(g_out[sl], dtype=dt)
            g_v = g * st                                          # d exp(v), v=-dbv*I (E folded: st = exp(v)*E)
            g_dbv[sl] = (g_v * (-I)).sum(-1)
            g_I = g_v * (-dbv[sl][:, None])
            # d I / d arg_k = W_k * (A_k/D_k) * j1(arg_k)   (bessel_j0' = -bessel_j1)
            coef = (W * (A / D)).astype(dt)
            j1 = _j1(arg, dt).astype(dt)
            g_arg = g_I[:, :, None] * coef[None, None, :] * j1
            g_t = (g_arg * (dt(1.5) * u)[None, None, :]).sum(-1, dtype=dt)
            g_dw = (g_t * taus[None, :]).sum(-1, dtype=dt)
            g_oef[sl] = g_dw * Ks
            g_hct[sl] = g_dw * (K0 * oef[sl])
    if want_grad:
        return (out, g_oef, g_dbv, g_hct) if want_hct else (out, g_oef, g_dbv)
    return out


def calc_blood(ph: Physics, oef, hct=None, dtype=F32, want_grad=False, g_out=None, want_hct=False):
    """calc_blood, live branch (signals.py:233-247)."""
    dt = dtype
    hct = ph.hct if hct is None else hct
    oef = np.asarray(oef, dtype=dt).reshape(-1)
    B, td = blood_b_of_tau(ph, dt)
    r2b = 1.0 / 0.189
    c0 = (4 / 45) * hct * (1 - hct)                               # Python double (or per-voxel array)
    c1 = 4.0 * math.pi * ph.b0 * ph.dchi
    c0 = np.asarray(c0, dtype=dt).reshape(-1) if np.ndim(c0) > 0 else dt(c0)
    base = dt(c1) * oef
    g0 = c0 * base ** 2                                           # :239
    G = (dt(0.5 * (ph.gamma ** 2)) * g0) * dt(td ** 2)            # :241
    Eb = _exp(-r2b * ph.te, dt)
    ex = np.exp(-G[:, None] * B[None, :]).astype(dt)
    sb = Eb * ex
    if not want_grad:
        return sb
    g = np.asarray(g_out, dtype=dt)
    g_G = (g * sb * (-B[None, :])).sum(-1)
    g_g0 = g_G * dt(0.5 * (ph.gamma ** 2)) * dt(td ** 2)
    g_base = g_g0 * c0 * dt(2.0) * base
    if want_hct:                                                  # c0 = (4/45) hct (1 - hct)
        g_hct = (g_g0 * base ** 2) * (dt(4 / 45) * (dt(1.0) - dt(2.0) * np.asarray(hct, dtype=dt).reshape(-1)))
        return sb, g_base * dt(c1), g_hct
    return sb, g_base * dt(c1)


def forward(ph: Physics, oef_dbv, full_model=True, include_blood=True, dtype=F32,
            variable_hct=False, chunk=4096):
    """SignalGenerationLayer.call without noise/misalignment (signals.py:55-114,137-140)."""
    x = np.asarray(oef_dbv)
    lead = x.shape[:-1]
    width = 3 if variable_hct else 2
    assert x.shape[-1] == width, 'Input should have %d elements in last dimension' % width
    flat = x.reshape(-1, width).astype(dtype)
    oef, dbv = flat[:, 0], flat[:, 1]
    hct = flat[:, 2] if variable_hct else None
    st = calc_tissue(ph, oef, dbv, hct, full_model, dtype, chunk)
    if include_blood:
        bw = (m_bld(ph, dtype) * dtype(NB)) * dbv                  # :107
        sb = calc_blood(ph, oef, hct, dtype)
    else:
        bw = dbv                                                  # :110
        sb = np.zeros_like(st)
    tw = dtype(1.0) - bw                                          # :112
    sig = tw[:, None] * st + bw[:, None] * sb                     # :114
    return sig.reshape(lead + (ph.n_tau,))


def forward_backward(ph: Physics, oef_dbv, g_signal, full_model=True, include_blood=True,
                     dtype=F32, chunk=4096, variable_hct=False):
    """Forward plus the TF-autodiff-consistent vector-Jacobian product.

    Returns (S [N,n_tau], g_oef_dbv [N,2]) for upstream gradient g_signal [N,n_tau];
    with ``variable_hct`` rows are (OEF, DBV, Hct) and the gradient is [N,3] (signals.py:64-70).
    Chain (what tape.gradient would do for signals.py:98-114): mix -> tissue / blood."""
    dt = dtype
    if variable_hct:
        flat = np.asarray(oef_dbv).reshape(-1, 3).astype(dt)
        g = np.asarray(g_signal).reshape(-1, ph.n_tau).astype(dt)
        oef, dbv, hct = flat[:, 0], flat[:, 1], flat[:, 2]
        kappa = m_bld(ph, dt) * dt(NB) if include_blood else dt(1.0)
        bw = kappa * dbv
        tw = dt(1.0) - bw
        st, go_t, gd_t, gh_t = calc_tissue(ph, oef, dbv, hct, full_model, dt, chunk, True, g * tw[:, None], True)
        if include_blood:
            sb, go_b, gh_b = calc_blood(ph, oef, hct, dt, True, g * bw[:, None], True)
        else:
            sb, go_b, gh_b = np.zeros_like(st), np.zeros_like(oef), np.zeros_like(oef)
        sig = tw[:, None] * st + bw[:, None] * sb
        g_dbv = gd_t + kappa * (g * (sb - st)).sum(-1)
        return sig, np.stack([go_t + go_b, g_dbv, gh_t + gh_b], -1).astype(dt)
    flat = np.asarray(oef_dbv).reshape(-1, 2).astype(dt)
    g = np.asarray(g_signal).reshape(-1, ph.n_tau).astype(dt)
    oef, dbv = flat[:, 0], flat[:, 1]
    if include_blood:
        kappa = m_bld(ph, dt) * dt(NB)
        bw = kappa * dbv
    else:
        kappa = dt(1.0)
        bw = dbv
    tw = dt(1.0) - bw
    g_st = g * tw[:, None]
    g_sb = g * bw[:, None]
    st, go_t, gd_t = calc_tissue(ph, oef, dbv, None, full_model, dt, chunk, True, g_st)
    if include_blood:
        sb, go_b = calc_blood(ph, oef, None, dt, True, g_sb)
    else:
        sb, go_b = np.zeros_like(st), np.zeros_like(oef)
    sig = tw[:, None] * st + bw[:, None] * sb
    g_bw = (g * (sb - st)).sum(-1)
    g_dbv = gd_t + kappa * g_bw
    g_oef = go_t + go_b
    return sig, np.stack([g_oef, g_dbv], -1).astype(dt)


def forward_misaligned(ph: Physics, oef_dbv, prob, sel_u01, from_index, eps, full_model=True, include_blood=True,
                       dtype=F32, variable_hct=False):
    """SignalGenerationLayer.call with misaligned_prob > 0, noise off (signals.py:80-96), explicit draws:
    sel_u01 [N] (uniform of :82), from_index [N] (randint of :84-85), eps [N,2] (the normals of :92-93).
    The reference turns OEF/DBV into per-image [N,n_tau] tensors; each image's signal depends on its own pair only,
    so the result is the per-image selection between the forward model of the original and the perturbed pair."""
    dt = dtype
    x = np.asarray(oef_dbv).astype(dt)
    x = x.reshape(-1, x.shape[-1])
    base = forward(ph, x, full_model, include_blood, dt, variable_hct)
    mis = np.asarray(sel_u01, dtype=dt).reshape(-1) < dt(prob)                                    # :82
    late = (np.arange(ph.n_tau)[None, :] > np.asarray(from_index).reshape(-1, 1)) & mis[:, None]  # :86-88
    e = np.asarray(eps, dtype=dt).reshape(-1, 2)
    pert = x.copy()
    pert[:, 0] = np.clip(e[:, 0] * dt(0.15) + x[:, 0], dt(0.05), dt(0.8))                        # :92
    pert[:, 1] = np.clip(e[:, 1] * dt(0.05) + x[:, 1], dt(0.002), dt(0.3))                       # :93
    alt = forward(ph, pert, full_model, include_blood, dt, variable_hct)
    return np.where(late, alt, base)                                                              # :95-96


def add_noise(signal, snr_u, eps, dtype=F32):
    """Noise model (signals.py:116-128) with explicit draws: snr_u ~ U(50,120) [N,1],
    eps ~ N(0,1) [N,n_tau].  The std uses the *batch* mean of the signal (:126)."""
    dt = dtype
    sig = np.asarray(signal, dtype=dt)
    nt = sig.shape[-1]
    if nt == 11:
        norm_snr = NORM_SNR_11.astype(dt)
    elif nt == 24:
        norm_snr = (1.0 - (np.abs(np.arange(-0.028, 0.065, 0.004)) * 3.0)).astype(dt)   # :121
    else:
        raise UnboundLocalError("norm_snr is only defined for 11 or 24 taus (signals.py:117-121)")
    snr = np.asarray(snr_u, dtype=dt).reshape(-1, 1) * norm_snr.reshape(1, nt)
    std = sig.mean(0, keepdims=True, dtype=dt) / snr
    return (sig + np.asarray(eps, dtype=dt) * std).astype(dt)


def calculate_r2p(ph: Physics, oef, dbv, hct=None, dtype=F32):
    """signals.py:149-150 / model.py:516-525."""
    K = dtype(dw_const(ph, hct))
    return (K * np.asarray(oef, dtype=dtype)) * np.asarray(dbv, dtype=dtype)


# --------------------------------------------------------------------------
# sampling / transforms (model.py:21-50, 288-316)
# --------------------------------------------------------------------------
def _sigmoid(x, dt):
    x = np.asarray(x, dtype=dt)
    return (dt(1.0) / (dt(1.0) + np.exp(-x))).astype(dt)


def transform_std(raw, dt=F32):
    return (np.tanh(np.asarray(raw, dtype=dt)) * dt(3.0)) - dt(1.0)      # model.py:288-290


def transform_offdiag(raw, dt=F32):
    return np.tanh(np.asarray(raw, dtype=dt)) * dt(np.exp(-2.0))         # model.py:292-294


def forward_transform(z, dt=F32):
    z = np.asarray(z, dtype=dt)
    oef = _sigmoid(z[..., 0], dt) * dt(OEF_RANGE) + dt(MIN_OEF)          # model.py:302
    dbv = _sigmoid(z[..., 1], dt) * dt(DBV_RANGE) + dt(MIN_DBV)          # model.py:303
    return np.stack([oef, dbv], -1)


def backwards_transform(signal, include_logit, dt=F32):
    s = np.asarray(signal, dtype=dt)
    oef = (s[..., 0] - dt(MIN_OEF)) / dt(OEF_RANGE)                      # model.py:310
    dbv = (s[..., 1] - dt(MIN_DBV)) / dt(DBV_RANGE)                      # model.py:311
    if include_logit:
        oef, dbv = logit(oef, dt), logit(dbv, dt)
    return np.stack([oef, dbv], -1)


def logit(x, dt=F32):
    x = np.asarray(x, dtype=dt)
    return np.log(x / (dt(1.0) - x)).astype(dt)                          # model.py:10-12


def reparam_sample(q, eps, use_mvg=True, dt=F32):
    """ReparamTrickLayer.call (model.py:21-50).  q [...,5|4], eps [...,2] -> (oef_dbv [...,2], z [...,2])."""
    q = np.asarray(q, dtype=dt)
    eps = np.asarray(eps, dtype=dt)
    sd_o = np.exp(transform_std(q[..., 1], dt))
    sd_d = np.exp(transform_std(q[..., 3], dt))
    z_o = q[..., 0] + eps[..., 0] * sd_o
    if use_mvg:
        z_d = (q[..., 2] + eps[..., 0] * transform_offdiag(q[..., 4], dt)) + eps[..., 1] * sd_d
    else:
        z_d = q[..., 2] + eps[..., 1] * sd_d
    z = np.stack([z_o, z_d], -1).astype(dt)
    return forward_transform(z, dt), z


# --------------------------------------------------------------------------
# likelihood (model.py:527-568)
# --------------------------------------------------------------------------
def fine_tune_nll(y_true, pred, sigma, se_idx, dt=F32, multi_image_normalisation=False,
                  predict_log_data=False, student_t_df=None, return_mean=True):
    """fine_tune_loss_fn (model.py:527-568), heteroscedastic branch.

    y_true [N, n_tau+1] (last channel = mask), pred [N, n_tau], sigma [N, n_tau]."""
    y_true = np.asarray(y_true, dtype=dt)
    pred = np.asarray(pred, dtype=dt)
    sigma = np.asarray(sigma, dtype=dt)
    nt = pred.shape[-1]
    mask = y_true[:, -1:]
    if multi_image_normalisation:
        yn = y_true / (y_true[:, se_idx - 1:se_idx + 2].mean(-1, keepdims=True, dtype=dt) + dt(1e-3))
        pn = pred / (pred[:, se_idx - 1:se_idx + 2].mean(-1, keepdims=True, dtype=dt) + dt(1e-3))
    else:
        yn = y_true / (y_true[:, se_idx:se_idx + 1] + dt(1e-3))          # :544
        pn = pred / (pred[:, se_idx:se_idx + 1] + dt(1e-3))              # :545
    if predict_log_data:
        with np.errstate(divide="ignore", invalid="ignore"):
            yn = np.where(mask > 0, np.log(yn), dt(0.0))
            pn = np.where(mask > 0, np.log(pn), dt(0.0))
    res = yn[:, :nt] - pn                                               # :552
    if student_t_df is not None and student_t_df < 50:
        df = float(student_t_df)
        zq = res / sigma
        logp = (math.lgamma(0.5 * (df + 1.0)) - math.lgamma(0.5 * df) - 0.5 * math.log(df * math.pi))
        nll = -(dt(logp) - np.log(sigma) - dt(0.5 * (df + 1.0)) * np.log1p(zq * zq / dt(df)))
    else:
        nll = -(-np.log(sigma) - dt(np.log(np.sqrt(2.0 * np.pi))) - dt(0.5) * np.square(res / sigma))  # :561
    nll = nll.sum(-1, keepdims=True, dtype=dt) * mask                   # :563-564
    if return_mean:
        return dt(nll.sum(dtype=dt) / mask.sum(dtype=dt))               # :566
    return nll


# --------------------------------------------------------------------------
# logit-normal log-prob and KL (model.py:376-447, 592-665)
# --------------------------------------------------------------------------
def logit_gaussian_mvg_nll(observations, params, dt=F32):
    """logit_gaussian_mvg_log_prob (model.py:376-400): returns the *negative* log prob
    (the reference's naming notwithstanding).  observations [N,2] in OEF/DBV units;
    params [N,5] raw."""
    obs = np.asarray(observations, dtype=dt)
    p = np.asarray(params, dtype=dt).reshape(-1, 5)
    mu_o, ls_o = p[:, 0], transform_std(p[:, 1], dt)
    mu_d, ls_d = p[:, 2], transform_std(p[:, 3], dt)
    cov = transform_offdiag(p[:, 4], dt)
    x = backwards_transform(obs, False, dt)
    x = np.clip(x, dt(1e-6), dt(1.0) - dt(1e-6))                        # :394-395
    zh = logit(x, dt)
    log_det = dt(2.0) * (ls_o + ls_d)                                   # :443-447
    inv_tl = np.exp(ls_o * dt(-1.0))                                    # :432
    inv_br = np.exp(ls_d * dt(-1.0))
    inv_bl = np.exp(ls_o * dt(-1.0) + ls_d * dt(-1.0)) * cov * dt(-1.0)
    r_o, r_d = zh[:, 0] - mu_o, zh[:, 1] - mu_d
    w_o = r_o * inv_tl
    w_d = r_d * inv_br + r_o * inv_bl
    swr = np.square(w_o) + np.square(w_d)
    loss = -(-dt(np.log(2.0 * np.pi)) - dt(0.5) * log_det - dt(0.5) * swr)   # :390
    loss = loss + (np.log(x) + np.log(dt(1.0) - x)).sum(-1, dtype=dt)   # :398
    return loss.astype(dt)


def mc_kl(prior, pred, eps_kl, dt=F32):
    """mvg_kl_samples (model.py:592-610).  prior/pred [N,5] raw, eps_kl [N,S,2].
    Returns per-voxel KL estimate [N]."""
    pred = np.asarray(pred, dtype=dt)
    prior = np.asarray(prior, dtype=dt)
    eps_kl = np.asarray(eps_kl, dtype=dt)
    S = eps_kl.shape[1]
    acc = []
    for s in range(S):
        smp, _ = reparam_sample(pred, eps_kl[:, s, :], True, dt)
        log_q = -logit_gaussian_mvg_nll(smp, pred, dt)
        log_p = -logit_gaussian_mvg_nll(smp, prior, dt)
        acc.append(log_q - log_p)
    return np.stack(acc, -1).mean(-1, dtype=dt).astype(dt)               # :609


def closed_form_kl(prior, pred, dt=F64):
    """Textbook KL( N(mu_q, L_q L_q^T) || N(mu_p, L_p L_p^T) ) in logit space, L=[[e^ls_o,0],[c,e^ls_d]].
    This is what mvg_kl (model.py:612-652) intends; the reference's version has a
    transposed-inverse slip in the trace term and is unused."""
    q = np.asarray(pred, dtype=dt)
    p = np.asarray(prior, dtype=dt)
    a_q, b_q, c_q = np.exp(transform_std(q[:, 1], dt)), np.exp(transform_std(q[:, 3], dt)), transform_offdiag(q[:, 4], dt)
    a_p, b_p, c_p = np.exp(transform_std(p[:, 1], dt)), np.exp(transform_std(p[:, 3], dt)), transform_offdiag(p[:, 4], dt)
    # M = L_p^{-1} L_q (lower triangular); trace(Sigma_p^{-1} Sigma_q) = ||M||_F^2
    m00 = a_q / a_p
    m10 = (c_q - c_p * m00) / b_p
    m11 = b_q / b_p
    tr = m00 ** 2 + m10 ** 2 + m11 ** 2
    d_o, d_d = p[:, 0] - q[:, 0], p[:, 2] - q[:, 2]
    w_o = d_o / a_p
    w_d = (d_d - c_p * w_o) / b_p
    maha = w_o ** 2 + w_d ** 2
    logdet = 2.0 * (np.log(a_p) + np.log(b_p) - np.log(a_q) - np.log(b_q))
    return (0.5 * (tr + maha - 2.0 + logdet)).astype(dt)


# --------------------------------------------------------------------------
# fused ELBO value + gradients (what tape.gradient gives for train.py:315-320 minus TV)
# --------------------------------------------------------------------------
def elbo_and_grads(ph: Physics, q, sigma, y, mask, prior, eps, eps_kl, dt=F32,
                   full_model=True, include_blood=True, se_idx=2, kl_weight=1.0,
                   multi_image_normalisation=False, chunk=2048):
    """NLL (fine_tune_loss_fn) + kl_weight * KL (kl_loss) for one batch and their
    gradients w.r.t. the encoder outputs q [N,5] and sigma [N,n_tau].

    Follows build_fine_tuner (model.py:239-286): sample -> forward model -> loss.
    Gradients are TF-autodiff-consistent: stop_gradient on q inside log q
    (model.py:596), identity gradient through the clip (model.py:395), bessel_j0'=-j1.
    Returns dict(nll, kl, grad_q, grad_sigma, nll_map, kl_map, grad_q_nll, grad_q_kl)."""
    q = np.asarray(q, dtype=dt).reshape(-1, 5)
    n = q.shape[0]
    nt = ph.n_tau
    sigma = np.asarray(sigma, dtype=dt).reshape(n, nt)
    y = np.asarray(y, dtype=dt).reshape(n, nt)
    mask = np.asarray(mask, dtype=dt).reshape(n)
    prior = np.asarray(prior, dtype=dt).reshape(n, 5)
    eps = np.asarray(eps, dtype=dt).reshape(n, 2)
    msum = mask.sum(dtype=dt)

    # ---- transforms and their local derivatives
    th1, th3, th4 = np.tanh(q[:, 1]), np.tanh(q[:, 3]), np.tanh(q[:, 4])
    ls_o, ls_d = th1 * dt(3.0) - dt(1.0), th3 * dt(3.0) - dt(1.0)
    cov = th4 * dt(np.exp(-2.0))
    sd_o, sd_d = np.exp(ls_o), np.exp(ls_d)
    dls_o, dls_d = dt(3.0) * (dt(1.0) - th1 * th1), dt(3.0) * (dt(1.0) - th3 * th3)
    dcov = dt(np.exp(-2.0)) * (dt(1.0) - th4 * th4)

    def sample(e):
        z_o = q[:, 0] + e[:, 0] * sd_o
        z_d = (q[:, 2] + e[:, 0] * cov) + e[:, 1] * sd_d
        s_o, s_d = _sigmoid(z_o, dt), _sigmoid(z_d, dt)
        oef = s_o * dt(OEF_RANGE) + dt(MIN_OEF)
        dbv = s_d * dt(DBV_RANGE) + dt(MIN_DBV)
        return oef, dbv, s_o, s_d

    def z_grads_to_q(gz_o, gz_d, e):
        gq = np.zeros((n, 5), dtype=dt)
        gq[:, 0] = gz_o
        gq[:, 1] = gz_o * e[:, 0] * sd_o * dls_o
        gq[:, 2] = gz_d
        gq[:, 3] = gz_d * e[:, 1] * sd_d * dls_d
        gq[:, 4] = gz_d * e[:, 0] * dcov
        return gq

    # ---- likelihood term
    oef, dbv, s_o, s_d = sample(eps)
    pred = forward(ph, np.stack([oef, dbv], -1), full_model, include_blood, dt, chunk=chunk)
    if multi_image_normalisation:
        ny = y[:, se_idx - 1:se_idx + 2].mean(-1, dtype=dt) + dt(1e-3)
        npd = pred[:, se_idx - 1:se_idx + 2].mean(-1, dtype=dt) + dt(1e-3)
    else:
        ny = y[:, se_idx] + dt(1e-3)
        npd = pred[:, se_idx] + dt(1e-3)
    yn = y / ny[:, None]
    pn = pred / npd[:, None]
    res = yn - pn
    nll_pix = (-(-np.log(sigma) - dt(np.log(np.sqrt(2.0 * np.pi))) - dt(0.5) * np.square(res / sigma))).sum(-1, dtype=dt)
    nll_map = nll_pix * mask
    nll = nll_map.sum(dtype=dt) / msum
    scale = mask / msum                                              # d nll / d nll_pix
    g_sigma = (dt(1.0) / sigma - np.square(res) / sigma ** 3) * scale[:, None]
    g_pn = (-(res) / np.square(sigma)) * scale[:, None]
    # pn_k = pred_k / npd  (npd depends on pred[se] or the 3-image mean)
    g_pred = g_pn / npd[:, None]
    g_npd = -(g_pn * pred).sum(-1, dtype=dt) / np.square(npd)
    if multi_image_normalisation:
        g_pred[:, se_idx - 1:se_idx + 2] += (g_npd / dt(3.0))[:, None]
    else:
        g_pred[:, se_idx] += g_npd
    _, g_od = forward_backward(ph, np.stack([oef, dbv], -1), g_pred, full_model, include_blood, dt, chunk=chunk)
    gz_o = g_od[:, 0] * dt(OEF_RANGE) * s_o * (dt(1.0) - s_o)
    gz_d = g_od[:, 1] * dt(DBV_RANGE) * s_d * (dt(1.0) - s_d)
    grad_q = z_grads_to_q(gz_o, gz_d, eps)

    # ---- KL term (70 samples in the reference: model.py:654)
    kl_map = np.zeros(n, dtype=dt)
    grad_q_kl = np.zeros((n, 5), dtype=dt)
    if eps_kl is not None:
        eps_kl = np.asarray(eps_kl, dtype=dt)
        S = eps_kl.shape[1]
        pls_o, pls_d = transform_std(prior[:, 1], dt), transform_std(prior[:, 3], dt)
        pcov = transform_offdiag(prior[:, 4], dt)

        def whitened(zh_o, zh_d, mu_o, mu_d, l_o, l_d, c):
            inv_tl, inv_br = np.exp(l_o * dt(-1.0)), np.exp(l_d * dt(-1.0))
            inv_bl = np.exp(l_o * dt(-1.0) + l_d * dt(-1.0)) * c * dt(-1.0)
            r_o, r_d = zh_o - mu_o, zh_d - mu_d
            w_o = r_o * inv_tl
            w_d = r_d * inv_br + r_o * inv_bl
            nll_ = -(-dt(np.log(2.0 * np.pi)) - dt(0.5) * (dt(2.0) * (l_o + l_d)) - dt(0.5) * (np.square(w_o) + np.square(w_d)))
            # d nll / d zh
            g_o = w_o * inv_tl + w_d * inv_bl
            g_d = w_d * inv_br
            return nll_, g_o, g_d

        for s in range(S):
            e = eps_kl[:, s, :]
            o_s, d_s, so_s, sd_s = sample(e)
            x_o = (o_s - dt(MIN_OEF)) / dt(OEF_RANGE)
            x_d = (d_s - dt(MIN_DBV)) / dt(DBV_RANGE)
            x_o = np.clip(x_o, dt(1e-6), dt(1.0) - dt(1e-6))
            x_d = np.clip(x_d, dt(1e-6), dt(1.0) - dt(1e-6))
            zh_o, zh_d = logit(x_o, dt), logit(x_d, dt)
            jac = (np.log(x_o) + np.log(dt(1.0) - x_o)) + (np.log(x_d) + np.log(dt(1.0) - x_d))
            nq, gq_o, gq_d = whitened(zh_o, zh_d, q[:, 0], q[:, 2], ls_o, ls_d, cov)
            npri, gp_o, gp_d = whitened(zh_o, zh_d, prior[:, 0], prior[:, 2], pls_o, pls_d, pcov)
            kl_map += (-(nq + jac)) - (-(npri + jac))                # log_q - log_p (:596-603)
            # d(log_q - log_p)/d zh = -(gq) + gp ; then zh -> x -> sample -> z
            gzh_o, gzh_d = gp_o - gq_o, gp_d - gq_d
            dzh_o = (dt(1.0) / (x_o * (dt(1.0) - x_o))) * so_s * (dt(1.0) - so_s)
            dzh_d = (dt(1.0) / (x_d * (dt(1.0) - x_d))) * sd_s * (dt(1.0) - sd_s)
            grad_q_kl += z_grads_to_q(gzh_o * dzh_o, gzh_d * dzh_d, e)
        kl_map = kl_map / dt(S)
        grad_q_kl = grad_q_kl / dt(S)
        live = mask > 0
        kl_map = np.where(live, kl_map, dt(0.0))                     # model.py:661
        grad_q_kl = np.where(live[:, None], grad_q_kl, dt(0.0)) / msum
    kl = kl_map.sum(dtype=dt) / msum
    return dict(nll=dt(nll), kl=dt(kl), elbo=dt(nll + dt(kl_weight) * kl),
                grad_q=(grad_q + dt(kl_weight) * grad_q_kl).astype(dt), grad_sigma=g_sigma.astype(dt),
                nll_map=nll_map, kl_map=kl_map, pred=pred,
                grad_q_nll=grad_q.astype(dt), grad_q_kl=grad_q_kl.astype(dt))


# --------------------------------------------------------------------------
# posterior statistics (model.py:318-343)
# --------------------------------------------------------------------------
def posterior_stats(ph: Physics, q, eps_s, dt=F32):
    """calculate_means(..., include_r2p=True, return_stds=True): eps_s [N,S,2].
    Returns means [N,3], 'stds' [N,3] (which are variances, model.py:331,337)."""
    q = np.asarray(q, dtype=dt).reshape(-1, 5)
    eps_s = np.asarray(eps_s, dtype=dt)
    S = eps_s.shape[1]
    smp = np.stack([reparam_sample(q, eps_s[:, s, :], True, dt)[0] for s in range(S)], -1)   # [N,2,S]
    means = smp.mean(-1, dtype=dt)
    var = np.square(smp - means[..., None]).mean(-1, dtype=dt)
    r2p = calculate_r2p(ph, smp[:, 0, :], smp[:, 1, :], None, dt)
    r2p_m = r2p.mean(-1, keepdims=True, dtype=dt)
    r2p_v = np.square(r2p - r2p_m).mean(-1, keepdims=True, dtype=dt)
    return np.concatenate([means, r2p_m], -1).astype(dt), np.concatenate([var, r2p_v], -1).astype(dt)


# --------------------------------------------------------------------------
# synthetic data generation (signals.py:251-300) with explicit draws
# --------------------------------------------------------------------------
def synthetic_dataset_from_draws(ph: Physics, oefs, dbvs, perm, snr_u=None, noise_eps=None,
                                 full_model=True, use_blood=True, dt=F32, n_chunks=10, misalign=None,
                                 variable_hct=False):
    """create_synthetic_dataset after the random draws: meshgrid ('ij') -> [constant Hct column 0.34 with
    variable_hct, signals.py:273-276] -> shuffle (explicit permutation) -> 10 chunked layer calls (misalignment
    :80-96 with ``misalign`` = (prob, sel_u01 [n], from_index [n], eps [n,2]) concatenated over the chunks, then the
    noise, whose std is a per-chunk statistic :126,282-285) -> labels [OEF, DBV, R2']."""
    oefs = np.asarray(oefs, dtype=dt)
    dbvs = np.asarray(dbvs, dtype=dt)
    xx, yy = np.meshgrid(oefs, dbvs, indexing='ij')                  # :270
    train_y = np.stack([xx.reshape(-1), yy.reshape(-1)], 1)
    if variable_hct:
        train_y = np.concatenate([train_y, np.full((train_y.shape[0], 1), 0.34, dtype=dt)], -1)   # :273-276
    train_y = train_y[np.asarray(perm)]                              # :279
    n = train_y.shape[0]
    chunk = n // n_chunks                                            # :283
    xs = []
    for i in range(n_chunks):
        sl = slice(i * chunk, (i + 1) * chunk)
        if misalign is not None:
            prob, mu, mi, me = misalign
            s = forward_misaligned(ph, train_y[sl], prob, mu[sl], mi[sl], me[sl], full_model, use_blood, dt,
                                   variable_hct)
        else:
            s = forward(ph, train_y[sl], full_model, use_blood, dt, variable_hct)
        if ph.simulate_noise:
            s = add_noise(s, snr_u[sl], noise_eps[sl], dt)
        xs.append(s)
    train_x = np.concatenate(xs, 0)                                  # rows beyond 10*chunk are dropped (:283-287)
    if variable_hct:
        r2p = (dt(dw_const(ph, 1.0)) * train_y[:, 2] * train_y[:, 0]) * train_y[:, 1]             # :293-294
    else:
        r2p = calculate_r2p(ph, train_y[:, 0], train_y[:, 1], None, dt)  # :296
    return train_x, np.concatenate([train_y[:, :2], r2p[:, None]], -1).astype(dt)


# --------------------------------------------------------------------------
# losses either side of the path
# --------------------------------------------------------------------------
def smoothness_loss(q, mask, dt=F64):
    """Total-variation term, model.py:726-754: x/y neighbour differences of the forward-transformed, range-rescaled
    means, both voxels inside the mask, sum|d| / sum(mask).  q [B,X,Y,Z,C>=4] (channels 0 and 2 are the means),
    mask [B,X,Y,Z].  Returns (value, d value / d q) -- sub-gradient sign(0) = 0 as tf.abs."""
    q = np.asarray(q, dtype=dt)
    mask = np.asarray(mask, dtype=dt)
    z = np.stack([q[..., 0], q[..., 2]], -1)
    s = _sigmoid(z, dt)
    rng = np.array([0.8, 0.2], dtype=dt)
    mn = np.array([0.04, 0.001], dtype=dt)
    p = (s * rng + mn) / rng                                                     # model.py:736-738
    g_p = np.zeros_like(p)
    total = dt(0)
    for axis in (1, 2):
        lo = [slice(None)] * 5
        hi = [slice(None)] * 5
        lo[axis], hi[axis] = slice(None, -1), slice(1, None)
        lo, hi = tuple(lo), tuple(hi)
        both = ((mask[lo[:4]] > 0) & (mask[hi[:4]] > 0))[..., None]
        d = np.where(both, p[lo] - p[hi], 0)                                     # model.py:741-747
        total = total + np.abs(d).sum()
        g_p[lo] += np.sign(d)
        g_p[hi] -= np.sign(d)
    msum = mask.sum()
    g = np.zeros_like(q)
    dz = g_p * (s * (1 - s)) / msum                                              # d p / d z = range * s(1-s) / range
    g[..., 0], g[..., 2] = dz[..., 0], dz[..., 1]
    return dt(total / msum), g


def _lgamma(x):
    import math as _m
    return _m.lgamma(x)


def synthetic_data_nll(labels, pred, use_mvg=True, inv_gamma_alpha=0.0, inv_gamma_beta=0.0, dt=F64):
    """Per-row pre-training loss of synthetic_data_loss (model.py:449-514, without the sampled r2p term) and its
    gradient w.r.t. the raw predictions.  labels [N,>=2] (OEF, DBV), pred [N,5] (mvg) or [N,4] (diagonal).
    The reference's result is mean(rows) (for the inverse-gamma term: see oracle/make_golden.py, section F)."""
    import math as _m
    y = np.asarray(labels, dtype=dt)
    q = np.asarray(pred, dtype=dt)
    x = backwards_transform(y[:, :2], False, dt)                                 # model.py:393 / :416
    th1, th3 = np.tanh(q[:, 1]), np.tanh(q[:, 3])
    ls_o, ls_d = th1 * 3 - 1, th3 * 3 - 1
    g = np.zeros_like(q)
    if use_mvg:
        x = np.clip(x, 1e-6, 1 - 1e-6)                                           # model.py:394-395
        th4 = np.tanh(q[:, 4])
        cov = th4 * np.exp(dt(-2.0))
        const = _m.log(2.0 * _m.pi)
    else:
        cov = np.zeros_like(ls_o)
        const = 0.0                                                              # gaussian_nll drops it (model.py:404)
    zl = np.log(x / (1 - x))
    r_o, r_d = zl[:, 0] - q[:, 0], zl[:, 1] - q[:, 2]
    inv_o, inv_d, e_neg = np.exp(-ls_o), np.exp(-ls_d), np.exp(-ls_o - ls_d)
    w_o = r_o * inv_o
    w_d = r_d * inv_d - r_o * e_neg * cov                                        # model.py:432-439
    loss = const + (ls_o + ls_d) + 0.5 * (w_o ** 2 + w_d ** 2)
    loss = loss + np.log(x[:, 0]) + np.log(1 - x[:, 0]) + np.log(x[:, 1]) + np.log(1 - x[:, 1])
    g[:, 0] = -(w_o * inv_o - w_d * e_neg * cov)
    g[:, 2] = -w_d * inv_d
    d_ls_o = 1 - w_o ** 2 + w_d * r_o * e_neg * cov
    d_ls_d = 1 - w_d ** 2
    d_cov = -w_d * r_o * e_neg
    d_q4 = np.zeros_like(ls_o)
    if inv_gamma_alpha * inv_gamma_beta > 0.0:                                   # model.py:495-507
        a, b = inv_gamma_alpha, inv_gamma_beta
        if use_mvg:
            v_o, v_d = np.exp(ls_o) ** 2, np.exp(ls_d) ** 2 + q[:, 4] ** 2       # raw channel 4, as the reference
            dv_d_ls, d_q4 = 2 * np.exp(ls_d) ** 2, 2 * q[:, 4]
        else:
            v_o, v_d = np.exp(ls_o * 2), np.exp(ls_d * 2)
            dv_d_ls = 2 * v_d
        lp = lambda v: a * _m.log(b) - _lgamma(a) - (a + 1) * np.log(v) - b / v  # noqa: E731
        dlp = lambda v: -(a + 1) / v + b / v ** 2                                # noqa: E731
        loss = loss - (lp(v_o) + lp(v_d))
        d_ls_o = d_ls_o - dlp(v_o) * 2 * v_o
        d_ls_d = d_ls_d - dlp(v_d) * dv_d_ls
        d_q4 = -dlp(v_d) * d_q4
    g[:, 1] = d_ls_o * 3 * (1 - th1 ** 2)
    g[:, 3] = d_ls_d * 3 * (1 - th3 ** 2)
    if use_mvg:
        g[:, 4] = d_cov * np.exp(dt(-2.0)) * (1 - th4 ** 2) + d_q4
    return loss.astype(dt), g.astype(dt)


def diag_kl(prior, pred, dt=F64):
    """KL of the diagonal (use_mvg=False) branch, model.py:685-708: tfp LogitNormal.kl_divergence = KL of the
    underlying Normals, summed over OEF and DBV.  prior / pred [N,4] raw.  Returns (kl, d/d pred, d/d prior)."""
    q, p = np.asarray(pred, dtype=dt), np.asarray(prior, dtype=dt)
    kl = np.zeros(q.shape[0], dtype=dt)
    gq, gp = np.zeros_like(q), np.zeros_like(p)
    for m, s in ((0, 1), (2, 3)):
        thq, thp = np.tanh(q[:, s]), np.tanh(p[:, s])
        lq, lp = thq * 3 - 1, thp * 3 - 1
        d = (q[:, m] - p[:, m]) * np.exp(-lp)
        kl += 0.5 * d ** 2 + 0.5 * np.expm1(2 * (lq - lp)) - (lq - lp)
        gq[:, m] = d * np.exp(-lp)
        gp[:, m] = -d * np.exp(-lp)
        gq[:, s] = (np.exp(2 * (lq - lp)) - 1) * 3 * (1 - thq ** 2)
        gp[:, s] = (-d ** 2 - np.exp(2 * (lq - lp)) + 1) * 3 * (1 - thp ** 2)
    return kl, gq, gp
