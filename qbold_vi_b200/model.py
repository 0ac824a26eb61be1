"""Host-side mirror of the sampling / likelihood / KL slice of the reference's ``model.py``.

Reference names and signatures are kept (``ReparamTrickLayer``, ``EncoderTrainer`` with
``transform_std``, ``forward_transform``, ``fine_tune_loss_fn(y_true, y_pred[, return_mean])``,
``kl_loss(true, predicted[, return_mean, no_samples])``, ``calculate_means`` ...), tensors are torch
CUDA tensors in the reference's channels-last layout ``[B, X, Y, Z, C]``.

The training hot path is ``EncoderTrainer.fused_elbo`` -- one sm_100a kernel for
sample -> forward model -> NLL (+ MC KL) -> backward (what ``build_fine_tuner`` + the two Keras loss
closures of train.py:315-320 + autodiff do in the reference).  The stand-alone loss callables exist
for drop-in parity of the API; small element-wise helpers (transforms, TV smoothness, the
pre-training NLL) are torch ops and run wherever their tensors live.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from ._lib import on_device as _on_device

from . import _lib
from ._lib import QboldLikelihood, QboldParams, check, dptr, stream_ptr
from .signals import SignalGenerationLayer


def logit(signal):
    return torch.log(signal / (1.0 - signal))                                  # model.py:10-12


def _as_mvg(params, use_mvg):
    """Diagonal posterior (use_mvg=False, 4 channels: model.py:33-37, 406-421, 695-708) == the Cholesky
    parameterisation with a zero raw off-diagonal (tanh(0) = 0): pad a fifth channel so the same kernels serve it."""
    if use_mvg:
        return params
    return torch.cat([params[..., :4], torch.zeros_like(params[..., :1])], -1)


def _check_mvg_operands(pred, prior):
    """The kernels take raw pointers: refuse anything but [...,5] operands with equal voxel counts (a reference-style
    10-channel 'predictions' tensor with an appended population prior would otherwise be read as 2N rows)."""
    if pred.shape[-1] != 5 or prior.shape[-1] != 5:
        raise ValueError('mvg KL: predicted and prior must have 5 channels (mean, raw std, mean, raw std, raw '
                         'off-diagonal); got %d and %d' % (pred.shape[-1], prior.shape[-1]))
    if pred.numel() != prior.numel():
        raise ValueError('mvg KL: %d predicted voxels but %d prior voxels' % (pred.numel() // 5, prior.numel() // 5))


def _next_seed(obj):
    obj._calls += 1
    return (obj._seed + 0x9E3779B97F4A7C15 * obj._calls) & 0xFFFFFFFFFFFFFFFF


class _ReparamFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, eps):
        n = q.shape[0]
        out = torch.empty((n, 2), dtype=torch.float32, device=q.device)
        with _on_device(q.device):
            check(_lib.lib().qbold_reparam_sample(dptr(q), dptr(eps), 0, 0, n, dptr(out), stream_ptr(q.device)))
        ctx.save_for_backward(q, eps, out)
        return out

    @staticmethod
    def backward(ctx, g):
        q, eps, out = ctx.saved_tensors
        th1, th3, th4 = torch.tanh(q[:, 1]), torch.tanh(q[:, 3]), torch.tanh(q[:, 4])
        sd_o, sd_d = torch.exp(th1 * 3.0 - 1.0), torch.exp(th3 * 3.0 - 1.0)
        s_o, s_d = (out[:, 0] - 0.04) / 0.8, (out[:, 1] - 0.001) / 0.2
        gz_o = g[:, 0] * 0.8 * s_o * (1.0 - s_o)
        gz_d = g[:, 1] * 0.2 * s_d * (1.0 - s_d)
        gq = torch.stack([gz_o, gz_o * eps[:, 0] * sd_o * 3.0 * (1.0 - th1 * th1), gz_d,
                          gz_d * eps[:, 1] * sd_d * 3.0 * (1.0 - th3 * th3),
                          gz_d * eps[:, 0] * math.exp(-2.0) * (1.0 - th4 * th4)], -1)
        return gq, None


class ReparamTrickLayer:
    """Draw samples of OEF and DBV from the predicted distributions (model.py:15-50)."""

    def __init__(self, encoder_trainer):
        self._encoder_trainer = encoder_trainer

    def __call__(self, inputs, *args, **kwargs):
        return self.call(inputs, *args, **kwargs)

    def call(self, inputs, *args, eps=None, **kwargs):
        input, mask = inputs
        lead = tuple(input.shape[:-1])
        q = _as_mvg(input, self._encoder_trainer._use_mvg).reshape(-1, 5).float().contiguous()
        if eps is None:
            eps = torch.randn((q.shape[0], 2), dtype=torch.float32, device=q.device)
        eps = eps.reshape(-1, 2).float().contiguous()
        return _ReparamFn.apply(q, eps).reshape(lead + (2,))


class _KlFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, prior, mask, eps_kl, seed, n_samples, offset=0):
        n = pred.shape[0]
        kl_map = torch.empty(n, dtype=torch.float32, device=pred.device)
        grad = torch.empty((n, 5), dtype=torch.float32, device=pred.device)
        with _on_device(pred.device):
            check(_lib.lib().qbold_kl(dptr(pred), dptr(prior), dptr(mask, allow_none=True),
                                      dptr(eps_kl, allow_none=True), seed, int(offset), n_samples, n, dptr(kl_map),
                                      dptr(grad),
                                      stream_ptr(pred.device)))
        ctx.save_for_backward(grad)
        return kl_map

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g[:, None], None, None, None, None, None, None


class _NllFn(torch.autograd.Function):
    """fine_tune_loss_fn's per-voxel map from predictions that already exist (qbold_nll); the kernel also returns
    the partial derivatives, so backward is two multiplies."""

    @staticmethod
    def forward(ctx, pred, sigma, y, mask, params):
        n, nt = pred.shape
        nll = torch.empty(n, dtype=torch.float32, device=pred.device)
        d_pred, d_sigma = torch.empty_like(pred), torch.empty_like(sigma)
        with _on_device(pred.device):
            check(_lib.lib().qbold_nll(C.byref(params), dptr(y), dptr(pred), dptr(sigma), dptr(mask), n, dptr(nll),
                                       dptr(d_pred), dptr(d_sigma), stream_ptr(pred.device)))
        ctx.save_for_backward(d_pred, d_sigma)
        return nll

    @staticmethod
    def backward(ctx, g):
        d_pred, d_sigma = ctx.saved_tensors
        return d_pred * g[:, None], d_sigma * g[:, None], None, None, None


class _TvFn(torch.autograd.Function):
    """smoothness_loss (qbold_smoothness): value and gradient from one pass over the encoder output."""

    @staticmethod
    def forward(ctx, q, mask, mask_sum):
        q = q.float().contiguous()
        b, x, y, z, c = q.shape
        grad = torch.empty_like(q)
        tv = torch.zeros(1, dtype=torch.float64, device=q.device)
        with _on_device(q.device):
            if torch.is_tensor(mask_sum):                       # global sum(mask) left on the device: no host sync
                inv = (1.0 / mask_sum.reshape(1).to(q.device, torch.float64)).float().contiguous()
                check(_lib.lib().qbold_smoothness_dev(dptr(q), c, dptr(mask), b, x, y, z, dptr(inv),
                                                      dptr(tv, torch.float64), dptr(grad), stream_ptr(q.device)))
                ctx.save_for_backward(grad)
                return (tv[0] * inv[0].double()).float()
            check(_lib.lib().qbold_smoothness(dptr(q), c, dptr(mask), b, x, y, z, 1.0 / mask_sum,
                                              dptr(tv, torch.float64), dptr(grad), stream_ptr(q.device)))
        ctx.save_for_backward(grad)
        return (tv[0] / mask_sum).float()

    @staticmethod
    def backward(ctx, g):
        return ctx.saved_tensors[0] * g, None, None


class _SynthNllFn(torch.autograd.Function):
    """mean over rows of the pre-training NLL (qbold_synth_nll)."""

    @staticmethod
    def forward(ctx, pred, labels, use_mvg, ig_alpha, ig_beta):
        n = pred.shape[0]
        grad = torch.empty_like(pred)
        total = torch.zeros(1, dtype=torch.float64, device=pred.device)
        with _on_device(pred.device):
            check(_lib.lib().qbold_synth_nll(dptr(labels), labels.shape[1], dptr(pred), int(use_mvg), float(ig_alpha),
                                             float(ig_beta), n, 1.0 / n, None, dptr(grad),
                                             dptr(total, torch.float64), stream_ptr(pred.device)))
        ctx.save_for_backward(grad)
        return (total[0] / n).float()

    @staticmethod
    def backward(ctx, g):
        return ctx.saved_tensors[0] * g, None, None, None, None


class _DiagKlFn(torch.autograd.Function):
    """Per-voxel KL of the diagonal branch (qbold_diag_kl).  ``both`` [n,8]: q and a trainable population prior side
    by side (model.py:687-689); otherwise pred [n,4] with a fixed prior [n,4]."""

    @staticmethod
    def forward(ctx, pred, prior, mask):
        n, width = pred.shape
        kl = torch.empty(n, dtype=torch.float32, device=pred.device)
        grad = torch.empty_like(pred)
        with _on_device(pred.device):
            if prior is None:                                   # population prior inside `pred`
                p, g = pred.data_ptr(), grad.data_ptr()
                check(_lib.lib().qbold_diag_kl(p, 8, p + 16, 8, dptr(mask, allow_none=True), n, dptr(kl), g, 8, g + 16,
                                               8, stream_ptr(pred.device)))
            else:
                check(_lib.lib().qbold_diag_kl(dptr(pred), 4, dptr(prior), 4, dptr(mask, allow_none=True), n, dptr(kl),
                                               dptr(grad), 4, None, 0, stream_ptr(pred.device)))
        ctx.save_for_backward(grad)
        return kl

    @staticmethod
    def backward(ctx, g):
        return ctx.saved_tensors[0] * g[:, None], None, None


class _MogKlFn(torch.autograd.Function):
    """Per-voxel single-sample KL against the mixture-of-Gaussians population prior (qbold_mog_kl)."""

    @staticmethod
    def forward(ctx, pred, mask, eps, seed, n_comp, offset):
        n = pred.shape[0]
        kl = torch.empty(n, dtype=torch.float32, device=pred.device)
        grad = torch.empty_like(pred)
        with _on_device(pred.device):
            check(_lib.lib().qbold_mog_kl(dptr(pred), n_comp, dptr(mask, allow_none=True), dptr(eps, allow_none=True),
                                          seed, int(offset), n, dptr(kl), dptr(grad), stream_ptr(pred.device)))
        ctx.save_for_backward(grad)
        return kl

    @staticmethod
    def backward(ctx, g):
        return ctx.saved_tensors[0] * g[:, None], None, None, None, None, None


class _SynthNllInferredFn(torch.autograd.Function):
    """Pre-training NLL with LEARNED InverseGamma parameters (qbold_synth_nll_inferred): gradients for the q channels
    come from the kernel, those of the 4 hyper-parameters from its 4 reduced sums."""

    @staticmethod
    def forward(ctx, pred, ig, labels, use_mvg):
        n, dev = pred.shape[0], pred.device
        nc = 5 if use_mvg else 4
        grad = torch.empty((n, nc), dtype=torch.float32, device=dev)
        acc = torch.zeros(5, dtype=torch.float64, device=dev)                 # loss sum | 4 hyper-parameter sums
        with _on_device(dev):
            check(_lib.lib().qbold_synth_nll_inferred(dptr(labels), labels.shape[1], dptr(pred), pred.shape[1],
                                                      int(use_mvg), dptr(ig), n, 1.0 / n, None, dptr(grad),
                                                      C.c_void_p(acc.data_ptr()), C.c_void_p(acc.data_ptr() + 8),
                                                      stream_ptr(dev)))
        a, b = ig.double()[0::2], ig.double()[1::2]
        g_alpha = -((torch.log(b) - torch.digamma(a)) - acc[1::2][:2] / n)    # d mean(loss) / d alpha_(oef, dbv)
        g_beta = -(a / b - acc[2::2][:2] / n)
        ctx.save_for_backward(grad, torch.stack([g_alpha[0], g_beta[0], g_alpha[1], g_beta[1]]).float())
        ctx.width = pred.shape[1]
        return (acc[0] / n).float()

    @staticmethod
    def backward(ctx, g):
        grad, g_ig = ctx.saved_tensors
        gp = torch.zeros((grad.shape[0], ctx.width), dtype=torch.float32, device=grad.device)
        gp[:, :grad.shape[1]] = grad * g
        return gp, g_ig * g, None, None


class _FusedElboFn(torch.autograd.Function):
    """loss = nll + kl_weight * kl of one (local) batch; gradients come from the same launch."""

    @staticmethod
    def forward(ctx, q, sigma, trainer, layer, y, mask, prior, eps, eps_kl, seed, kl_samples, inv_mask_sum,
                kl_weight, want_maps, offset=0):
        n = q.shape[0]
        dev = q.device
        nt = layer.n_tau
        grad_q = torch.empty((n, 5), dtype=torch.float32, device=dev)
        grad_sigma = torch.empty((n, nt), dtype=torch.float32, device=dev)
        sums = torch.zeros(4, dtype=torch.float64, device=dev)
        nll_map = torch.empty(n, dtype=torch.float32, device=dev) if want_maps else None
        kl_map = torch.empty(n, dtype=torch.float32, device=dev) if want_maps else None
        with _on_device(dev):
            if torch.is_tensor(seed):       # Philox key on the device (int64 bit pattern of the uint64): captured steps
                if not torch.is_tensor(inv_mask_sum) or eps is not None or eps_kl is not None:
                    raise ValueError('fused_elbo: a device seed needs a device mask_sum and in-kernel draws (no eps)')
                check(_lib.lib().qbold_elbo_fused_graph(
                    C.byref(trainer._params_for(layer)), dptr(q), dptr(sigma), dptr(y), dptr(mask),
                    dptr(prior, allow_none=True), dptr(seed, torch.int64), int(offset), kl_samples, dptr(inv_mask_sum),
                    kl_weight, n, dptr(grad_q), dptr(grad_sigma), dptr(nll_map, allow_none=True),
                    dptr(kl_map, allow_none=True), dptr(sums, torch.float64), stream_ptr(dev)))
                fn = None
            else:
                fn = _lib.lib().qbold_elbo_fused_dev if torch.is_tensor(inv_mask_sum) else _lib.lib().qbold_elbo_fused
            if fn is not None:
                check(fn(
                    C.byref(trainer._params_for(layer)), dptr(q), dptr(sigma), dptr(y), dptr(mask),
                    dptr(prior, allow_none=True), dptr(eps, allow_none=True), dptr(eps_kl, allow_none=True), seed,
                    int(offset), kl_samples, dptr(inv_mask_sum) if torch.is_tensor(inv_mask_sum) else inv_mask_sum,
                    kl_weight, n, dptr(grad_q), dptr(grad_sigma),
                    dptr(nll_map, allow_none=True), dptr(kl_map, allow_none=True), dptr(sums, torch.float64),
                    stream_ptr(dev)))
        ctx.save_for_backward(grad_q, grad_sigma)
        s = sums.float()
        ims = inv_mask_sum[0] if torch.is_tensor(inv_mask_sum) else inv_mask_sum
        nll, kl = s[0] * ims, s[1] * ims
        ctx.mark_non_differentiable(sums)
        if want_maps:
            ctx.mark_non_differentiable(nll_map, kl_map)
            return nll + kl_weight * kl, sums, nll_map, kl_map
        return nll + kl_weight * kl, sums

    @staticmethod
    def backward(ctx, g, *unused):
        grad_q, grad_sigma = ctx.saved_tensors
        return (grad_q * g, grad_sigma * g) + (None,) * 13


class EncoderTrainer:
    """Reference model.py:53-95 (constructor signature preserved)."""

    def __init__(self, system_params, no_intermediate_layers=1, no_units=10, use_layer_norm=False, dropout_rate=0.0,
                 activation_type='gelu', student_t_df=None, initial_im_sigma=0.08, multi_image_normalisation=True,
                 channelwise_gating=False, infer_inv_gamma=False, use_mvg=True, use_population_prior=True,
                 mog_components=1, no_samples=1, heteroscedastic_noise=True, predict_log_data=True, seed=None):
        self._no_intermediate_layers = no_intermediate_layers
        self._no_units = no_units
        self._use_layer_norm = use_layer_norm
        self._dropout_rate = dropout_rate
        self._activation_type = activation_type
        self._student_t_df = student_t_df
        self._initial_im_sigma = initial_im_sigma
        self._multi_image_normalisation = multi_image_normalisation
        self._system_params = system_params
        self._channelwise_gating = channelwise_gating
        self._infer_inv_gamma = infer_inv_gamma
        self._use_mvg = use_mvg
        self._use_population_prior = use_population_prior
        self._mog_components = mog_components
        self._no_samples = no_samples
        self._oef_range = 0.8
        self._min_oef = 0.04
        self._dbv_range = 0.2
        self._min_dbv = 0.001
        self._heteroscedastic_noise = heteroscedastic_noise
        self._predict_log_data = predict_log_data
        self._se_idx = int(abs(float(system_params['tau_start']) / float(system_params['tau_step'])))   # model.py:95
        self._seed = int(torch.initial_seed() if seed is None else seed) & 0xFFFFFFFFFFFFFFFF
        self._calls = 0

    # ------------------------------------------------------------------ transforms (model.py:288-316)
    def transform_std(self, pred_stds):
        return (torch.tanh(pred_stds) * 3.0) - 1.0

    def transform_offdiag(self, pred_offdiag):
        return torch.tanh(pred_offdiag) * np.exp(-2.0)

    def inv_transform_std(self, std):
        return torch.atanh((std + 1.0) / 3.0)

    def forward_transform(self, logits):
        oef, dbv = torch.split(logits, 1, -1)
        oef = (torch.sigmoid(oef) * self._oef_range) + self._min_oef
        dbv = (torch.sigmoid(dbv) * self._dbv_range) + self._min_dbv
        return torch.cat([oef, dbv], -1)

    def backwards_transform(self, signal, include_logit):
        oef, dbv = torch.split(signal, 1, -1)
        oef = (oef - self._min_oef) / self._oef_range
        dbv = (dbv - self._min_dbv) / self._dbv_range
        if include_logit:
            oef, dbv = logit(oef), logit(dbv)
        return torch.cat([oef, dbv], -1)

    def calculate_dw(self, oef):
        sp = self._system_params
        return SignalGenerationLayer.calculate_dw_static(oef, float(sp['hct']), float(sp['gamma']), float(sp['b0']),
                                                         float(sp['dchi']))

    def calculate_r2p(self, oef, dbv):
        return self.calculate_dw(oef) * dbv

    # ------------------------------------------------------------------ parameter block with likelihood options
    def _params_for(self, layer):
        """The layer's parameter block with this trainer's likelihood options.  Cached ON the layer object (it dies with
        it -- an id()-keyed dict would hand a recycled id the block of a dead layer) under the current option values, so
        later changes of the trainer's flags are seen."""
        df = self._student_t_df
        key = (self._se_idx, bool(self._multi_image_normalisation), bool(self._predict_log_data),
               float(df) if df is not None else 0.0)
        cache = layer.__dict__.setdefault('_likelihood_blocks', {})
        if key not in cache:
            p = QboldParams()
            C.memmove(C.byref(p), C.byref(layer.params), C.sizeof(QboldParams))
            lik = QboldLikelihood(key[0], int(key[1]), int(key[2]), 0, key[3])
            check(_lib.lib().qbold_params_set_likelihood(C.byref(p), C.byref(lik)))
            cache[key] = p
        return cache[key]

    # ------------------------------------------------------------------ sampling helpers (model.py:318-343)
    def create_samples(self, predicted_params, mask, no_samples):
        rpl = ReparamTrickLayer(self)
        return torch.stack([rpl([predicted_params, mask]) for _ in range(no_samples)], -1)

    def calculate_means(self, predicted_params, mask, include_r2p=False, return_stds=False, no_samples=20, eps=None,
                        signal_layer=None, offset=0):
        """Posterior means (and the reference's 'stds', which are variances: model.py:331,337)."""
        lead = tuple(predicted_params.shape[:-1])
        q = _as_mvg(predicted_params, self._use_mvg).reshape(-1, 5).float().contiguous()
        n = q.shape[0]
        mean3 = torch.empty((n, 3), dtype=torch.float32, device=q.device)
        var3 = torch.empty((n, 3), dtype=torch.float32, device=q.device)
        layer = signal_layer or self._default_layer()
        e = None if eps is None else eps.reshape(n, no_samples, 2).float().contiguous()
        with _on_device(q.device):
            check(_lib.lib().qbold_posterior_stats(C.byref(layer.params), dptr(q), dptr(e, allow_none=True),
                                                   _next_seed(self), int(offset), no_samples, n, dptr(mean3), dptr(var3),
                                                   stream_ptr(q.device)))
        k = 3 if include_r2p else 2
        means = mean3[:, :k].reshape(lead + (k,))
        if return_stds:
            return means, var3[:, :k].reshape(lead + (k,))
        return means

    def _default_layer(self):
        if not hasattr(self, '_layer'):
            sp = dict(self._system_params)
            sp['simulate_noise'] = 'False'
            self._layer = SignalGenerationLayer(sp, True, True)
        return self._layer

    # ------------------------------------------------------------------ likelihood (model.py:527-568)
    def fine_tune_loss_fn(self, y_true, y_pred, return_mean=True, signal_layer=None):
        """model.py:527-568.  y_true [...,n_tau+1] (last channel = mask), y_pred [...,2*n_tau] (images, sigmas)."""
        y_true = torch.cat([y_true for _ in range(self._no_samples)], 0)
        no_images = y_true.shape[-1] - 1
        if not self._heteroscedastic_noise:
            return self._fine_tune_loss_homoscedastic(y_true, y_pred, return_mean)
        if y_pred.shape[-1] != 2 * no_images:
            raise ValueError('y_pred must hold %d predicted images followed by %d sigmas' % (no_images, no_images))
        layer = signal_layer or self._layer_for_tau_count(no_images)
        pred = y_pred[..., :no_images].reshape(-1, no_images).float().contiguous()
        sigma = y_pred[..., no_images:].reshape(-1, no_images).float().contiguous()
        y = y_true[..., :no_images].reshape(-1, no_images).float().contiguous()
        mask = y_true[..., -1].reshape(-1).float().contiguous()
        nll = _NllFn.apply(pred, sigma, y, mask, self._params_for(layer))
        if return_mean:
            return torch.sum(nll) / torch.sum(mask)
        return nll.reshape(-1, 1)

    def _layer_for_tau_count(self, no_images):
        layer = self._default_layer()
        if layer.n_tau != no_images:
            raise ValueError('fine_tune_loss_fn: %d images but the system parameters define %d taus; pass signal_layer='
                             % (no_images, layer.n_tau))
        return layer

    def _fine_tune_loss_homoscedastic(self, y_true, y_pred, return_mean):
        """heteroscedastic_noise=False branch (model.py:535-537): a single scalar sigma; plain tensor ops."""
        mask = y_true[..., -1:]
        no_images = y_true.shape[-1] - 1
        sigma = torch.mean(y_pred[..., -1:])
        y_pred = y_pred[..., :-1]
        se = self._se_idx
        if self._multi_image_normalisation:
            y_true = y_true / (torch.mean(y_true[..., se - 1:se + 2], -1, keepdim=True) + 1e-3)
            y_pred = y_pred / (torch.mean(y_pred[..., se - 1:se + 2], -1, keepdim=True) + 1e-3)
        else:
            y_true = y_true / (torch.mean(y_true[..., se:se + 1], -1, keepdim=True) + 1e-3)
            y_pred = y_pred / (torch.mean(y_pred[..., se:se + 1], -1, keepdim=True) + 1e-3)
        if self._predict_log_data:
            y_true = torch.where(mask > 0, torch.log(y_true), torch.zeros_like(y_true))
            y_pred = torch.where(mask > 0, torch.log(y_pred), torch.zeros_like(y_pred))
        residual = (y_true[..., :-1] - y_pred).reshape(-1, no_images)
        mask = mask.reshape(-1, 1)
        if self._student_t_df is not None and self._student_t_df < 50:
            df = float(self._student_t_df)
            c = math.lgamma(0.5 * (df + 1.0)) - math.lgamma(0.5 * df) - 0.5 * math.log(df * math.pi)
            zq = residual / sigma
            nll = -(c - torch.log(sigma) - 0.5 * (df + 1.0) * torch.log1p(zq * zq / df))
        else:
            nll = -(-torch.log(sigma) - np.log(np.sqrt(2.0 * np.pi)) - 0.5 * torch.square(residual / sigma))
        nll = torch.sum(nll, -1, keepdim=True) * mask
        if return_mean:
            return torch.sum(nll) / torch.sum(mask)
        return nll

    # ------------------------------------------------------------------ KL (model.py:592-665)
    def mvg_kl_samples(self, prior, pred, no_samples=50, eps=None):
        prior_dist, mask = prior[..., :5], prior[..., 5:6]
        _check_mvg_operands(pred, prior_dist)
        lead = tuple(pred.shape[:-1])
        kl = _KlFn.apply(pred.reshape(-1, 5).float().contiguous(), prior_dist.reshape(-1, 5).float().contiguous(),
                         None, None if eps is None else eps.reshape(-1, no_samples, 2).float().contiguous(),
                         _next_seed(self), no_samples)
        return kl.reshape(lead + (1,))

    def kl_loss(self, true, predicted, return_mean=True, no_samples=70, eps=None, offset=0):
        """KL(q || prior), model.py:654-724.  mvg: the reference's 70-sample MC estimator (``no_samples=0`` selects
        the closed form); diagonal: the analytic LogitNormal KL, optionally against a trainable population prior
        carried in channels 4..7 of ``predicted`` plus its InverseGamma(1, 2) hyper-prior (:710-715)."""
        true = torch.cat([true for _ in range(self._no_samples)], 0)
        if self._use_mvg:
            prior_dist, mask = true[..., :5], true[..., 5:6]
            _check_mvg_operands(predicted, prior_dist)
            lead = tuple(predicted.shape[:-1])
            m = mask.reshape(-1).float().contiguous()
            kl = _KlFn.apply(predicted.reshape(-1, 5).float().contiguous(),
                             prior_dist.reshape(-1, 5).float().contiguous(), m,
                             None if eps is None else eps.reshape(-1, no_samples, 2).float().contiguous(),
                             _next_seed(self), no_samples, int(offset))
            if return_mean:
                return torch.sum(kl) / torch.sum(mask)
            return kl.reshape(lead + (1,))
        mask = true[..., 4:5]
        if self._use_population_prior and self._mog_components > 1:
            return self._mog_kl(predicted, mask, return_mean, eps, offset)

        lead = tuple(predicted.shape[:-1])
        m = mask.reshape(-1).float().contiguous()
        prior_cost = 0.0
        if self._use_population_prior:
            pred8 = predicted.reshape(-1, 8).float().contiguous()
            kl = _DiagKlFn.apply(pred8, None, m)
            ls = self.transform_std(predicted[..., [5, 7]])                       # p_oef_log_std, p_dbv_log_std
            a, b = 1.0, 2.0                                                       # InverseGamma(1, 2), model.py:711

            def ig_log_prob(v):
                return a * math.log(b) - math.lgamma(a) - (a + 1.0) * torch.log(v) - b / v

            prior_cost = -ig_log_prob(torch.exp(torch.mean(ls[..., 1]) * 2.0))
            prior_cost = prior_cost - ig_log_prob(torch.exp(torch.mean(ls[..., 0]) * 2.0))
            prior_cost = prior_cost * float(predicted.shape[0])
        else:
            kl = _DiagKlFn.apply(predicted.reshape(-1, 4).float().contiguous(),
                                 true[..., :4].reshape(-1, 4).float().contiguous(), m)
        if return_mean:
            return (torch.sum(kl) + prior_cost) / torch.sum(mask)
        return kl.reshape(lead + (1,))

    def _mog_kl(self, predicted, mask, return_mean, eps=None, offset=0):
        """Mixture-of-Gaussians population prior (model.py:666-684) in one kernel (``qbold_mog_kl``): single-sample
        estimate -entropy(q) + mean over components of the Gaussian NLL of one logit-space draw, with its gradient for
        q and every component.  ``predicted`` [...,4*(M+1)]: q then M components; ``eps`` [...,2] pins the (OEF, DBV)
        draws (default: in-kernel Philox)."""
        m_comp = self._mog_components
        width = 4 * (m_comp + 1)
        if predicted.shape[-1] != width:
            raise ValueError('kl_loss: expected %d channels (q + %d mixture components), got %d'
                             % (width, m_comp, predicted.shape[-1]))
        lead = tuple(predicted.shape[:-1])
        pred = predicted.reshape(-1, width).float().contiguous()
        m = mask.reshape(-1).float().contiguous()
        if m.shape[0] != pred.shape[0]:
            raise ValueError('kl_loss: %d mask voxels for %d predicted voxels' % (m.shape[0], pred.shape[0]))
        e = None if eps is None else eps.reshape(-1, 2).float().contiguous()
        kl = _MogKlFn.apply(pred, m, e, _next_seed(self), m_comp, int(offset))
        if return_mean:
            return torch.sum(kl) / torch.sum(mask)
        return kl.reshape(lead + (1,))

    # ------------------------------------------------------------------ fused training objective
    def fused_elbo(self, signal_layer, q_params, im_sigma, data, mask, prior, kl_samples=70, kl_weight=1.0,
                   eps=None, eps_kl=None, mask_sum=None, seed=None, return_maps=False, offset=0):
        """nll + kl_weight * kl for one batch in ONE kernel launch (differentiable w.r.t. q_params, im_sigma).
        ``offset``: global index of this batch's first voxel -- the in-kernel Philox counter is the global voxel
        index, so shards of one batch draw exactly what the unsharded batch would (pass it when sharding over ranks).

        q_params [...,5], im_sigma [...,n_tau], data [...,n_tau] (pre-masked, train.py:56), mask [...,1],
        prior [...,5] raw or None.  ``mask_sum`` = global sum(mask) when the batch is sharded over ranks.
        Returns (loss, dict(nll, kl[, nll_map, kl_map])) with the reference's normalisation (sum / sum(mask))."""
        nt = signal_layer.n_tau
        if not self._use_mvg:
            q_params, kl_samples = _as_mvg(q_params, False), 0          # analytic KL, as the reference (model.py:695-708)
            prior = None if prior is None else _as_mvg(prior, False)
        if q_params.shape[-1] != 5 or (prior is not None and prior.shape[-1] != 5):
            raise ValueError('fused_elbo: q_params / prior must have %d channels' % (5 if self._use_mvg else 4))
        q = q_params.reshape(-1, 5).float().contiguous()
        n = q.shape[0]
        for name, t, width in (('im_sigma', im_sigma, nt), ('data', data, nt), ('mask', mask, 1), ('prior', prior, 5)):
            if t is not None and t.numel() != n * width:
                raise ValueError('fused_elbo: %s holds %d values, expected %d voxels x %d' % (name, t.numel(), n, width))
        sg = im_sigma.reshape(n, nt).float().contiguous()
        y = data.reshape(n, nt).float().contiguous()
        m = mask.reshape(n).float().contiguous()
        pr = None if prior is None else prior.reshape(n, 5).float().contiguous()
        e = None if eps is None else eps.reshape(n, 2).float().contiguous()
        ek = None if eps_kl is None else eps_kl.reshape(n, kl_samples, 2).float().contiguous()
        if mask_sum is None:
            mask_sum = m.sum(dtype=torch.float64)                # stays on the device: no host synchronisation
        if torch.is_tensor(mask_sum):
            inv = (1.0 / mask_sum.reshape(1).to(q.device, torch.float64)).float().contiguous()
        else:
            inv = 1.0 / float(mask_sum)
        out = _FusedElboFn.apply(q, sg, self, signal_layer, y, m, pr, e, ek,
                                 _next_seed(self) if seed is None else seed, kl_samples if pr is not None else 0,
                                 inv, float(kl_weight), return_maps, int(offset))
        loss, sums = out[0], out[1]
        ims = inv[0] if torch.is_tensor(inv) else inv
        info = {'nll': (sums[0] * ims).float(), 'kl': (sums[1] * ims).float(), 'mask_sum': sums[2],
                'non_finite': sums[3]}
        if return_maps:
            info['nll_map'], info['kl_map'] = out[2], out[3]
        return loss, info

    # ------------------------------------------------------------------ glue graph (model.py:239-286)
    def build_fine_tuner(self, encoder_model, signal_generation_layer, input_im=None, input_mask=None):
        """Returns the fine-tuning model: ``model(data, mask)`` gives the reference's output dict
        {'predictions': q [...,5], 'predicted_images': concat[signal, sigma] [...,2*n_tau]} through the separate
        layers (sample -> forward model), and ``model.fused_loss(...)`` the one-launch training objective."""
        return FineTuner(self, encoder_model, signal_generation_layer)

    # ------------------------------------------------------------------ whole-volume inference (model.py:772-887)
    def likelihood_map(self, signal_layer, q_params, im_sigma, data, mask=None, no_samples=100, eps=None, offset=0):
        """Average per-voxel NLL over ``no_samples`` stochastic forward passes (save_predictions, model.py:808-817)."""
        nt = signal_layer.n_tau
        lead = tuple(q_params.shape[:-1])
        q = q_params.reshape(-1, 5).float().contiguous()
        n = q.shape[0]
        sg = im_sigma.reshape(n, nt).float().contiguous()
        y = data.reshape(n, nt).float().contiguous()
        m = None if mask is None else mask.reshape(n).float().contiguous()
        e = None if eps is None else eps.reshape(n, no_samples, 2).float().contiguous()
        out = torch.empty(n, dtype=torch.float32, device=q.device)
        with _on_device(q.device):
            check(_lib.lib().qbold_nll_map(C.byref(self._params_for(signal_layer)), dptr(q), dptr(sg), dptr(y),
                                           dptr(m, allow_none=True), dptr(e, allow_none=True), _next_seed(self),
                                           int(offset),
                                           no_samples, n, dptr(out), stream_ptr(q.device)))
        return out.reshape(lead + (1,))

    def posterior_inference(self, signal_layer, q_params, im_sigma, data, mask, prior=None, no_samples=64, offset=0):
        """BASELINE config 4: per-voxel posterior summaries of a whole volume in three launches:
        mean / variance of OEF, DBV, R2' (calculate_means), likelihood map and KL map."""
        means, variances = self.calculate_means(q_params, mask, include_r2p=True, return_stds=True,
                                                no_samples=no_samples, signal_layer=signal_layer, offset=offset)
        out = {'means': means, 'variances': variances,
               'likelihood': self.likelihood_map(signal_layer, q_params, im_sigma, data, mask, no_samples,
                                                 offset=offset)}
        if prior is not None:
            out['kl'] = self.kl_loss(torch.cat([prior, mask], -1), q_params, return_mean=False, no_samples=no_samples,
                                     offset=offset)
        return out

    def save_predictions(self, model, data, filename, transform_directory=None, use_first_op=True,
                         fine_tuner_model=None, priors=None, affine=None):
        """model.py:772-887: posterior maps of whole volumes as NIfTI images -- ``filename``_oef / _dbv / _r2p (means of
        200 draws) and _logstds (their variances, as the reference), plus, when ``fine_tuner_model`` is given,
        _likelihood (mean NLL map of 100 stochastic forward passes), _kl (100-sample KL map against ``priors``) and
        _residual (mean absolute normalised residual of one pass).  ``data`` [S,X,Y,Z,n_tau+1], last channel = mask.
        The FSL warp to MNI space (``transform_directory``) shells out to external binaries and is not provided."""
        from .nifti import save_im_data
        if transform_directory is not None:
            raise NotImplementedError('save_predictions: the FSL applywarp / fslmerge step (model.py:846-877) needs '
                                      'external binaries and is not provided')
        dev = next(model.parameters()).device
        data = torch.as_tensor(np.asarray(data, dtype=np.float32) if not torch.is_tensor(data) else data).float().to(dev)
        images, mask = data[..., :-1].contiguous(), data[..., -1:].contiguous()
        with torch.no_grad():
            p1, p2, im_sigma = model(images * mask)
            predictions = p1 if use_first_op else p2
            ones = torch.ones_like(predictions[..., :1])
            means, variances = self.calculate_means(predictions, ones, include_r2p=True, return_stds=True,
                                                    no_samples=200)
            if fine_tuner_model is not None:
                layer = fine_tuner_model.layer
                _, q, sigma = fine_tuner_model.encoder(images)                  # model.py:810: unmasked images
                lik = self.likelihood_map(layer, q, sigma, images, mask, no_samples=100)
                save_im_data(lik.cpu().numpy(), filename + '_likelihood', affine)
                if priors is not None:
                    pri = torch.as_tensor(np.asarray(priors, dtype=np.float32) if not torch.is_tensor(priors)
                                          else priors).float().to(dev)
                    kl = self.kl_loss(torch.cat([pri, mask], -1), q, return_mean=False, no_samples=100)
                    save_im_data(kl.cpu().numpy(), filename + '_kl', affine)
                y_pred = fine_tuner_model(images, mask)['predicted_images'][..., :images.shape[-1]]
                se = self._se_idx
                sl = slice(se - 1, se + 2) if self._multi_image_normalisation else slice(se, se + 1)
                y_true = images / (images[..., sl].mean(-1, keepdim=True) + 1e-3)
                y_pred = y_pred / (y_pred[..., sl].mean(-1, keepdim=True) + 1e-3)
                residual = (y_true - y_pred).abs().mean(-1, keepdim=True)
                save_im_data(residual.cpu().numpy(), filename + '_residual', affine)
        means, variances = means.cpu().numpy(), variances.cpu().numpy()
        for i, name in enumerate(('_oef', '_dbv', '_r2p')):
            save_im_data(means[..., i:i + 1], filename + name, affine)
        save_im_data(variances, filename + '_logstds', affine)
        return {'means': means, 'variances': variances}

    # ------------------------------------------------------------------ losses either side of the path
    def smoothness_loss(self, true_params, pred_params, mask_sum=None):
        """Total-variation term (model.py:726-754): x/y neighbours of the forward-transformed means, one stencil
        kernel for value + gradient.  ``mask_sum``: global sum(mask) when the batch is sharded over ranks."""
        true_params = torch.cat([true_params for _ in range(self._no_samples)], 0)
        c = 5 if self._use_mvg else 4
        mask = true_params[..., c].float().contiguous()
        if pred_params.dim() != 5 or pred_params.shape[-1] != c:
            raise ValueError('smoothness_loss: pred_params must be [B,X,Y,Z,%d]' % c)
        if mask_sum is None:
            mask_sum = mask.sum(dtype=torch.float64)             # stays on the device: no host synchronisation
        return _TvFn.apply(pred_params, mask, mask_sum if torch.is_tensor(mask_sum) else float(mask_sum))

    def oef_dbv_metrics(self, y_true, y_pred, oef_dbv_r2p=0, eps=None):
        """MSE of the 20-sample posterior means against the labels (model.py:345-364)."""
        means = self.calculate_means(y_pred, None, include_r2p=True, eps=eps)
        residual = means.reshape(-1, 3) - y_true.reshape(-1, 3)
        return torch.mean(torch.square(residual[:, min(int(oef_dbv_r2p), 2)]))

    def oef_metric(self, y_true, y_pred, eps=None):
        return self.oef_dbv_metrics(y_true, y_pred, 0, eps)

    def dbv_metric(self, y_true, y_pred, eps=None):
        return self.oef_dbv_metrics(y_true, y_pred, 1, eps)

    def r2p_metric(self, y_true, y_pred, eps=None):
        return self.oef_dbv_metrics(y_true, y_pred, 2, eps)

    def logit_gaussian_mvg_log_prob(self, observations, predicted_params):
        """model.py:376-400 (returns the negative log prob, as the reference does)."""
        shape = tuple(predicted_params.shape[:-1])
        p = predicted_params.reshape(-1, 5)
        ls_o, ls_d = self.transform_std(p[:, 1]), self.transform_std(p[:, 3])
        cov = self.transform_offdiag(p[:, 4])
        x = self.backwards_transform(observations[:, 0:2], False)
        x = x + (torch.clamp(x, 1e-6, 1.0 - 1e-6) - x).detach()              # clip_by_value_preserve_gradient
        z = logit(x)
        r_o, r_d = z[:, 0] - p[:, 0], z[:, 1] - p[:, 2]
        w_o = r_o * torch.exp(-ls_o)
        w_d = r_d * torch.exp(-ls_d) - r_o * torch.exp(-ls_o - ls_d) * cov
        loss = math.log(2.0 * math.pi) + 0.5 * (2.0 * (ls_o + ls_d)) + 0.5 * (w_o ** 2 + w_d ** 2)
        loss = loss + torch.sum(torch.log(x) + torch.log(1.0 - x), -1)
        return loss.reshape(shape)

    @staticmethod
    def gaussian_nll(obs, mean, log_std):
        return -(-log_std - 0.5 * ((obs - mean) / torch.exp(log_std)) ** 2)                      # model.py:402-404

    def synthetic_data_loss(self, y_true_orig, y_pred_orig, use_r2p_loss=False, inv_gamma_alpha=0.0,
                            inv_gamma_beta=0.0, eps=None):
        """Pre-training loss (model.py:449-514): mean over label rows of the logit-normal NLL (mvg or diagonal) minus
        the optional InverseGamma log-prior of the predicted variances, in one kernel with its gradient.  The
        reference adds its per-row extra terms ([N]) to a [B,X,Y,Z] loss, which only broadcasts for [N,1,1,1]
        inputs and then averages to mean(nll) + mean(extra); that is what is returned here for any shape.
        ``use_r2p_loss``: Gaussian NLL of the R2' label under 10 reparameterised draws (:480-494; ``eps``
        [N,10,2] pins the draws)."""
        c = 5 if self._use_mvg else 4
        labels = y_true_orig.reshape(-1, 3).float().contiguous()
        if self._infer_inv_gamma:
            # model.py:454-455 splits the prediction into two equal halves [q | hyper-prior], which only exists for the
            # diagonal layout (4 + 4 channels); the hyper-parameters are read from voxel 0 (:494)
            if y_pred_orig.shape[-1] != 2 * c or self._use_mvg:
                raise ValueError('infer_inv_gamma needs an 8-channel diagonal prediction [q(4) | alpha_oef, beta_oef, '
                                 'alpha_dbv, beta_dbv] (tf.split(y_pred, 2, -1), model.py:455); got %d channels, use_mvg=%s'
                                 % (y_pred_orig.shape[-1], self._use_mvg))
            both = y_pred_orig.reshape(-1, 2 * c).float().contiguous()
            pred = both[:, :c]
            loss = _SynthNllInferredFn.apply(both, both[0, c:], labels, False)
        else:
            pred = y_pred_orig.reshape(-1, c).float().contiguous()
            loss = _SynthNllFn.apply(pred, labels, self._use_mvg, float(inv_gamma_alpha), float(inv_gamma_beta))
        if use_r2p_loss:
            n_samples = 10
            rpl = ReparamTrickLayer(self)
            draws = [rpl((pred, None), eps=None if eps is None else eps.reshape(-1, n_samples, 2)[:, i])
                     for i in range(n_samples)]
            draws = torch.stack(draws, -1)                                                       # [N,2,10]
            r2p = self.calculate_r2p(draws[:, 0, :], draws[:, 1, :])
            r2p_log_std = torch.log(torch.std(r2p, -1, unbiased=False))
            loss = loss + torch.mean(self.gaussian_nll(labels[:, 2], torch.mean(r2p, -1), r2p_log_std))
        return loss


class FineTuner(torch.nn.Module):
    """build_fine_tuner (model.py:239-286) as a module; the encoder is any callable returning
    (q_voxelwise, q_spatial, sigma) like qbold_vi_b200.encoder.Encoder.

    ``use_population_prior`` (diagonal layout): a trainable prior vector (model.py:252-271; [-0.97, 0.4, -1.14, 0.6], or
    N(0,1) draws for a mixture of ``mog_components`` Gaussians) is broadcast over the voxels and appended to
    'predictions', as the reference does, so ``kl_loss`` on that output trains it.  With ``use_mvg`` the reference appends
    5 more channels that its own mvg ``kl_loss`` then mis-splits (ReparamTrickLayer splits 10 channels into 5 pairs,
    model.py:24,594): that combination is refused here instead of being reproduced."""

    def __init__(self, trainer, encoder_model, signal_generation_layer):
        super().__init__()
        self.trainer, self.encoder, self.layer = trainer, encoder_model, signal_generation_layer
        self.reparam = ReparamTrickLayer(trainer)
        self.pop_prior = None
        if trainer._use_population_prior:
            if trainer._use_mvg:
                raise NotImplementedError(
                    'use_population_prior with use_mvg: the reference concatenates a 5-channel prior onto the predictions '
                    '(model.py:254-271) but its mvg kl_loss never reads it as a prior (model.py:594 samples from the '
                    '10-channel tensor instead); use use_mvg=False for a trainable population prior')
            m = trainer._mog_components
            init = torch.randn(4 * m) if m > 1 else torch.tensor([-0.97, 0.4, -1.14, 0.6])     # model.py:261-264
            self.pop_prior = torch.nn.Parameter(init.float())

    def _with_pop_prior(self, q):
        if self.pop_prior is None:
            return q
        return torch.cat([q, self.pop_prior.to(q.device).expand(q.shape[:-1] + (self.pop_prior.numel(),))], -1)

    def forward(self, data, mask, eps=None):
        _, q, sigma = self.encoder(data)
        k = self.trainer._no_samples
        q_rep, sigma_rep = torch.cat([q] * k, 0), torch.cat([sigma] * k, 0)               # model.py:245-246
        sampled = self.reparam((q_rep, mask), eps=eps)                                    # :248
        pred = self.layer(sampled)                                                        # :273
        return {'predictions': self._with_pop_prior(q_rep),                               # :268-271
                'predicted_images': torch.cat([pred, sigma_rep], -1)}                     # :276,284-285

    def fused_loss(self, data, mask, prior, **kw):
        """One-launch training objective.  With a population prior the KL runs against the trainable prior
        (``kl_loss`` on the appended channels, which carries its gradient) and the fused kernel does the likelihood."""
        _, q, sigma = self.encoder(data)
        if self.pop_prior is None:
            return self.trainer.fused_elbo(self.layer, q, sigma, data, mask, prior, **kw)
        kl_weight = float(kw.pop('kl_weight', 1.0))
        for k in ('kl_samples', 'eps_kl'):
            kw.pop(k, None)
        nll, info = self.trainer.fused_elbo(self.layer, q, sigma, data, mask, None, **kw)
        true = torch.cat([torch.zeros_like(q[..., :4]), mask.reshape(q.shape[:-1] + (1,))], -1)
        kl = self.trainer.kl_loss(true, self._with_pop_prior(q))
        info['kl'] = kl.detach()
        return nll + kl_weight * kl, info
