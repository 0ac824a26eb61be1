"""Host-side mirror of the reference's ``signals.py``: same names, same call signatures,
same parameter mapping -- the arithmetic runs in the sm_100a kernels of libqbold.so.

    SignalGenerationLayer(system_parameters, full_model, include_blood,
                          misaligned_prob=0.0, variable_hct=False)          signals.py:18
    layer(x)            x[..., 2|3] CUDA float32 -> [..., n_tau]             signals.py:55-140
    create_synthetic_dataset(params, full_model, use_blood, misaligned_prob,
                             variable_hct=False, uniform_prop=0.1)           signals.py:251-300

Tensors are torch CUDA tensors (the reference uses tf.Tensors); the layer is differentiable
through torch.autograd with the gradient TensorFlow autodiff would produce.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from ._lib import on_device as _on_device

from . import _lib
from ._lib import QboldError, QboldLikelihood, QboldParams, QboldPhysics, check, dptr, stream_ptr


def _as_bool(v):
    """The reference CLI passes the strings 'True'/'False' (signals.py:330); accept both."""
    if isinstance(v, str):
        if v not in ('True', 'False'):
            raise ValueError('Arguments must be a valid boolean')            # signals.py:327-328
        return v == 'True'
    return bool(v)


def make_taus(tau_start, tau_end, tau_step):
    """tf.range(start, limit, delta, dtype=float32) (signals.py:34-35): float32 element i = start + i*delta."""
    s, e, d = np.float32(tau_start), np.float32(tau_end), np.float32(tau_step)
    n = int(math.ceil(abs(float((e - s) / d))))
    return (s + np.arange(n, dtype=np.float32) * d).astype(np.float32)


class _ForwardFn(torch.autograd.Function):
    """signal = layer(x); backward = TF-autodiff-consistent VJP recomputed by the fused kernel
    (nothing but the 8-byte/voxel input is saved)."""

    @staticmethod
    def forward(ctx, flat, layer):
        ctx.layer = layer
        ctx.save_for_backward(flat)
        return layer._forward_raw(flat)

    @staticmethod
    def backward(ctx, g):
        (flat,) = ctx.saved_tensors
        _, grad = ctx.layer.forward_backward(flat, g.contiguous())          # [N,2], or [N,3] with variable_hct
        return grad, None


class SignalGenerationLayer:
    """Forward ASE qBOLD signal model (reference signals.py:13-248)."""

    def __init__(self, system_parameters, full_model, include_blood, misaligned_prob=0.0, variable_hct=False,
                 taus=None, seed=None):
        sp = system_parameters
        self._gamma = float(sp['gamma'])
        self._b0 = float(sp['b0'])
        self._dchi = float(sp['dchi'])
        self._te = float(sp['te'])
        self._r2t = float(sp['r2t'])
        if taus is None:
            taus = make_taus(float(sp['tau_start']), float(sp['tau_end']), float(sp['tau_step']))
        self._taus = np.ascontiguousarray(taus, dtype=np.float32)
        self._tr = float(sp['tr'])
        self._ti = float(sp['ti'])
        self._t1b = float(sp['t1b'])
        self._simulate_noise = sp['simulate_noise'] == 'True'                 # signals.py:41
        self._weighted_noise = sp['tau_weighted'] == 'True'                   # parsed, unused (as in the reference)
        self._snr = int(sp['snr'])                                            # parsed, unused (as in the reference)
        # the reference only sets .hct when not variable_hct (signals.py:45-46)
        self.hct = float(sp['hct'])
        self._full_model = _as_bool(full_model)
        self._include_blood = _as_bool(include_blood)
        self._misaligned_prob = float(misaligned_prob)
        self._variable_hct = bool(variable_hct)
        self._seed = int(torch.initial_seed() if seed is None else seed) & 0xFFFFFFFFFFFFFFFF
        self._calls = 0

        phys = QboldPhysics(self._gamma, self._b0, self._dchi, self._te, self._r2t, self._tr, self._ti, self._t1b,
                            self.hct)
        self.params = QboldParams()
        tau_arr = (C.c_float * len(self._taus))(*self._taus.tolist())
        check(_lib.lib().qbold_params_init(C.byref(self.params), C.byref(phys), tau_arr, len(self._taus),
                                           int(self._full_model), int(self._include_blood)))

    # ------------------------------------------------------------------ reference API
    @property
    def n_tau(self):
        return len(self._taus)

    def __call__(self, input, *args, **kwargs):
        return self.call(input, *args, **kwargs)

    def call(self, input, *args, **kwargs):
        if not torch.is_tensor(input) and (hasattr(input, '__dlpack__') or type(input).__name__ == 'PyCapsule'):
            input = torch.from_dlpack(input)                # zero-copy hand-off (tf.experimental.dlpack, CuPy, JAX, ...)
        width = 3 if self._variable_hct else 2
        if self._variable_hct:
            assert input.shape[-1] == 3, 'Input should have 3 elements in last dimension, OEF, DBV and hct'
        else:
            assert input.shape[-1] == 2, 'Input should have 2 elements in last dimension, OEF and DBV'
        if not input.is_cuda:
            raise QboldError('qbold_vi_b200 runs on CUDA tensors only (got a %s tensor); there is no CPU path'
                             % input.device)
        flat = input.reshape(-1, width)
        if flat.dtype != torch.float32:
            flat = flat.float()
        flat = flat.contiguous()
        if flat.requires_grad and torch.is_grad_enabled():
            signal = _ForwardFn.apply(flat, self)
        else:
            signal = self._forward_raw(flat)
        if self._misaligned_prob > 0.0:
            if signal.requires_grad:
                signal = self._misalign_autograd(flat, signal)
            else:
                signal = self.misalign(flat, signal, inplace=True)
        if self._simulate_noise:
            noisy = self.add_noise(signal)
            # the noise is additive with unit Jacobian (signals.py:128); the reference's std_dev also depends on the
            # batch-mean signal (:126) -- that second-order path is not followed
            signal = signal + (noisy - signal.detach()) if signal.requires_grad else noisy
        return signal.reshape(tuple(input.shape[:-1]) + (self.n_tau,))

    def misalign(self, oef_dbv, signal, sel_u01=None, from_index=None, eps=None, seed=None, offset=0, inplace=False):
        """Misalignment augmentation (signals.py:80-96) in one launch (``qbold_misalign``): a Bernoulli(p) subset of
        voxels gets, for the images after a random index in [4, n_tau-1), the signal of perturbed parameters
        (OEF + N(0,0.15) clipped to [0.05,0.8], DBV + N(0,0.05) clipped to [0.002,0.3]).  ``signal`` is the noise-free
        forward model of ``oef_dbv``.  ``sel_u01`` [N], ``from_index`` [N] int32 and ``eps`` [N,2] pin the reference's
        draws (uniform, randint, the two normals); by default they come from the in-kernel Philox stream."""
        flat = oef_dbv.reshape(-1, oef_dbv.shape[-1]).float().contiguous()
        n = flat.shape[0]
        sig = signal.detach().reshape(n, self.n_tau)
        sig = sig if (inplace and sig.is_contiguous()) else sig.clone().contiguous()
        if n == 0:
            return sig
        if seed is None:
            seed = (self._seed + 0x9E3779B97F4A7C15 * (self._calls + 1)) & 0xFFFFFFFFFFFFFFFF
            self._calls += 1
        fi = None if from_index is None else from_index.reshape(n).to(torch.int32).contiguous()
        with _on_device(flat.device):
            check(_lib.lib().qbold_misalign(C.byref(self.params), dptr(flat), flat.shape[1], n, self._misaligned_prob,
                                            dptr(None if sel_u01 is None else sel_u01.reshape(n).float().contiguous(),
                                                 allow_none=True),
                                            dptr(fi, torch.int32, allow_none=True),
                                            dptr(None if eps is None else eps.reshape(n, 2).float().contiguous(),
                                                 allow_none=True),
                                            seed, int(offset), dptr(sig), stream_ptr(flat.device)))
        return sig

    def _misalign_autograd(self, flat, signal):
        """The same augmentation as differentiable tensor ops (only taken when the input requires grad): gradients reach
        OEF/DBV through the unperturbed images and, inside the clip range, through the perturbed ones (:92-96)."""
        n, nt = flat.shape[0], self.n_tau
        dev = flat.device
        misaligned = torch.rand(n, device=dev) < self._misaligned_prob                               # :82
        from_index = torch.randint(4, nt - 1, (n,), device=dev)                                      # :84-85
        idx = torch.nonzero(misaligned).reshape(-1)
        if idx.numel() == 0:
            return signal
        sel = flat[idx]
        cols = [(torch.randn(idx.numel(), device=dev) * 0.15 + sel[:, 0]).clamp(0.05, 0.8),          # :92
                (torch.randn(idx.numel(), device=dev) * 0.05 + sel[:, 1]).clamp(0.002, 0.3)]         # :93
        if flat.shape[1] == 3:
            cols.append(sel[:, 2])
        s2 = _ForwardFn.apply(torch.stack(cols, -1).contiguous(), self)
        late = torch.arange(nt, device=dev)[None, :] > from_index[idx, None]                         # :86-88
        return signal.index_copy(0, idx, torch.where(late, s2, signal[idx]))                         # :95-96

    @staticmethod
    def calculate_dw_static(oef, hct, gamma, b0, dchi):
        return (4.0 / 3.0) * math.pi * gamma * b0 * dchi * hct * oef            # signals.py:142-144

    def calculate_dw(self, oef, hct):
        return SignalGenerationLayer.calculate_dw_static(oef, hct, self._gamma, self._b0, self._dchi)

    def calculate_r2p(self, oef, dbv, hct):
        return self.calculate_dw(oef, hct) * dbv

    # ------------------------------------------------------------------ kernels
    def _forward_raw(self, flat):
        n, width = flat.shape
        out = torch.empty((n, self.n_tau), dtype=torch.float32, device=flat.device)
        with _on_device(flat.device):
            check(_lib.lib().qbold_forward(C.byref(self.params), dptr(flat), width, n, dptr(out),
                                           stream_ptr(flat.device)))
        return out

    def forward_backward(self, oef_dbv, g_signal=None, want_signal=True):
        """Forward + VJP in one fused launch: returns (signal [N,n_tau] or None, grad [N,2]); with variable_hct the rows
        are (OEF, DBV, Hct) and grad is [N,3]."""
        if not oef_dbv.is_cuda:
            raise QboldError('qbold_vi_b200 runs on CUDA tensors only (got a %s tensor); there is no CPU path'
                             % oef_dbv.device)
        width = 3 if self._variable_hct else 2
        flat = oef_dbv.reshape(-1, width).contiguous()
        n = flat.shape[0]
        sig = torch.empty((n, self.n_tau), dtype=torch.float32, device=flat.device) if want_signal else None
        grad = torch.empty((n, width), dtype=torch.float32, device=flat.device)
        g = None if g_signal is None else g_signal.reshape(n, self.n_tau).contiguous()
        fn = _lib.lib().qbold_forward_backward_hct if self._variable_hct else _lib.lib().qbold_forward_backward
        with _on_device(flat.device):
            check(fn(C.byref(self.params), dptr(flat), dptr(g, allow_none=True), n, dptr(sig, allow_none=True),
                     dptr(grad), stream_ptr(flat.device)))
        return sig, grad

    def forward_backward_host(self, oef_dbv, g_signal, signal_out, grad_out):
        """Host-buffer variant (pinned CPU tensors in, pinned CPU tensors out): the end-to-end path."""
        for t in (oef_dbv, g_signal, signal_out, grad_out):
            if t is not None and (t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous()):
                raise QboldError('forward_backward_host expects contiguous float32 CPU tensors')
        n = oef_dbv.shape[0]
        ptr = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        check(_lib.lib().qbold_forward_backward_host(C.byref(self.params), ptr(oef_dbv), ptr(g_signal), n,
                                                     ptr(signal_out), ptr(grad_out)))
        return signal_out, grad_out

    def add_noise(self, signal, snr_u01=None, eps=None, seed=None, offset=0, inplace=False):
        """Noise model of signals.py:116-128 (batch-mean statistic + per-voxel SNR)."""
        n = signal.shape[0]
        if n == 0:
            return signal
        sig = signal.detach()
        sig = sig if (inplace and sig.is_contiguous()) else sig.clone().contiguous()
        dev = sig.device
        mean = torch.empty(self.n_tau, dtype=torch.float32, device=dev)
        scratch = torch.empty(2 * self.n_tau, dtype=torch.float64, device=dev)
        if seed is None:
            seed = (self._seed + 0x9E3779B97F4A7C15 * (self._calls + 1)) & 0xFFFFFFFFFFFFFFFF
            self._calls += 1
        with _on_device(dev):
            check(_lib.lib().qbold_column_mean(dptr(sig), n, self.n_tau, dptr(mean), dptr(scratch, torch.float64),
                                               stream_ptr(dev)))
            if self.params.norm_snr[0] == 0.0:
                raise UnboundLocalError("local variable 'norm_snr' referenced before assignment "
                                        "(only 11 or 24 taus are supported, signals.py:117-121)")
            check(_lib.lib().qbold_add_noise(C.byref(self.params), dptr(sig), n, dptr(mean),
                                             dptr(snr_u01, allow_none=True), dptr(eps, allow_none=True),
                                             seed, offset, stream_ptr(dev)))
        return sig


def _truncated_normal(n, loc, scale, low, high, generator, device):
    """tfp.distributions.TruncatedNormal(loc, scale, low, high).sample(n) (signals.py:265-267) by inverse CDF."""
    u = torch.rand(n, dtype=torch.float64, device=device, generator=generator)
    a, b = (low - loc) / scale, (high - loc) / scale
    cdf = lambda x: 0.5 * (1.0 + math.erf(x / math.sqrt(2.0)))
    p = cdf(a) + u * (cdf(b) - cdf(a))
    x = loc + scale * torch.special.ndtri(p)
    return x.clamp(low, high).float()


def create_synthetic_dataset(params, full_model, use_blood, misaligned_prob, variable_hct=False, uniform_prop=0.1,
                             device=None, seed=None, shuffle='perm'):
    """Reference signals.py:251-300 on one GPU.  Returns (train_x [10*(S^2//10), n_tau], train_y [S^2, 3]).

    ``shuffle='perm'`` materialises an explicit permutation (torch.randperm, as tf.random.shuffle does);
    ``shuffle='feistel'`` uses the in-kernel keyed bijection (no permutation array; large S)."""
    device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    sig_layer = SignalGenerationLayer(params, full_model, use_blood, misaligned_prob=misaligned_prob,
                                      variable_hct=variable_hct, seed=seed)
    gen = torch.Generator(device=device)
    gen.manual_seed(sig_layer._seed & 0x7FFFFFFFFFFFFFFF)
    S = int(params['sample_size'])
    n_u, n_n = round(S * uniform_prop), round(S * (1.0 - uniform_prop))
    o0, o1 = float(params['oef_start']), float(params['oef_end'])
    d0, d1 = float(params['dbv_start']), float(params['dbv_end'])
    oefs = torch.rand(n_u, device=device, generator=gen) * (o1 - o0) + o0                       # :255-256
    oefs_n = torch.randn(n_n, device=device, generator=gen) * float(params['oef_std']) + float(params['oef_mean'])
    oefs = torch.cat([oefs, oefs_n.clamp(o0, o1)], 0).contiguous()                               # :258-260
    dbvs = torch.rand(n_u, device=device, generator=gen) * (d1 - d0) + d0                        # :262-263
    dbvs_n = _truncated_normal(n_n, float(params['dbv_mean']), float(params['dbv_std']), d0, d1, gen, device)
    dbvs = torch.cat([dbvs, dbvs_n], 0).contiguous()                                             # :265-268
    total = oefs.numel() * dbvs.numel()
    perm = None
    if shuffle == 'perm':
        perm = torch.randperm(total, device=device, generator=gen).contiguous()                  # :279
    return generate_from_marginals(sig_layer, oefs, dbvs, perm, n_chunks=10)


def generate_from_marginals(sig_layer, oefs, dbvs, perm=None, n_chunks=10, snr_u01=None, noise_eps=None,
                            seed=None, misalign_u01=None, misalign_index=None, misalign_eps=None):
    """signals.py:270-299 after the random draws.  Per chunk the reference calls the layer once: misalignment
    (if the layer's misaligned_prob > 0, signals.py:80-96) on the clean signal, then the noise, whose std is the
    per-call batch mean (signals.py:126, 282-285).  ``misalign_*``: the reference's recorded draws, concatenated
    over the chunks ([n_x], [n_x] int, [n_x,2]); default = the in-kernel Philox stream (chunk-invariant).
    With ``variable_hct`` every row carries the constant Hct 0.34 the reference draws (signals.py:273-276)."""
    lib = _lib.lib()
    dev = oefs.device
    nt = sig_layer.n_tau
    total = oefs.numel() * dbvs.numel()
    chunk = total // n_chunks                                                                    # :283
    n_x = chunk * n_chunks                                            # trailing rows are dropped from x only (:283-287)
    seed = sig_layer._seed if seed is None else seed
    train_x = torch.empty((n_x, nt), dtype=torch.float32, device=dev)
    train_y = torch.empty((total, 3), dtype=torch.float32, device=dev)
    pptr = None if perm is None else dptr(perm, torch.int64)
    with _on_device(dev):
        st = stream_ptr(dev)
        if sig_layer._variable_hct:                                 # labels only; the signals follow below
            check(lib.qbold_generate(C.byref(sig_layer.params), dptr(oefs), oefs.numel(), dptr(dbvs), dbvs.numel(),
                                     pptr, seed, 0, total, None, dptr(train_y), st))
        elif n_x > 0:               # rows are independent: the reference's chunking (:282-285) only matters for the noise
            check(lib.qbold_generate(C.byref(sig_layer.params), dptr(oefs), oefs.numel(), dptr(dbvs), dbvs.numel(),
                                     pptr, seed, 0, n_x, dptr(train_x), dptr(train_y), st))
        if total > n_x and not sig_layer._variable_hct:
            check(lib.qbold_generate(C.byref(sig_layer.params), dptr(oefs), oefs.numel(), dptr(dbvs), dbvs.numel(),
                                     pptr, seed, n_x, total - n_x, None, dptr(train_y[n_x:]), st))
    if sig_layer._variable_hct and total > 0:
        # rows (OEF, DBV, 0.34) through the variable-Hct forward kernel; R2' label with the float32 Hct (:293-294)
        hct = torch.full((total, 1), 0.34, dtype=torch.float32, device=dev)
        rows = torch.cat([train_y[:, :2], hct], -1).contiguous()
        if n_x > 0:
            train_x = sig_layer._forward_raw(rows[:n_x].contiguous())
        k = (4.0 / 3.0) * math.pi * sig_layer._gamma * sig_layer._b0 * sig_layer._dchi
        train_y[:, 2] = (k * rows[:, 2] * rows[:, 0]) * rows[:, 1]
    if sig_layer._misaligned_prob > 0.0 and n_x > 0:
        rows = rows[:n_x] if sig_layer._variable_hct else train_y[:n_x, :2].contiguous()
        train_x = sig_layer.misalign(rows, train_x, misalign_u01, misalign_index, misalign_eps,
                                     seed=(seed ^ 0x2545F4914F6CDD1D) & 0xFFFFFFFFFFFFFFFF, inplace=True)
    if sig_layer._simulate_noise and chunk > 0:
        # every chunk is one forward call in the reference, so its noise std uses that chunk's column means (:126)
        if sig_layer.params.norm_snr[0] == 0.0:
            raise UnboundLocalError("local variable 'norm_snr' referenced before assignment "
                                    "(only 11 or 24 taus are supported, signals.py:117-121)")
        scratch = torch.empty(32 * n_chunks, dtype=torch.float64, device=dev)
        with _on_device(dev):
            check(lib.qbold_add_noise_chunked(C.byref(sig_layer.params), dptr(train_x), chunk, n_chunks,
                                              dptr(None if snr_u01 is None else snr_u01[:n_x].contiguous(),
                                                   allow_none=True),
                                              dptr(None if noise_eps is None else noise_eps[:n_x].contiguous(),
                                                   allow_none=True),
                                              (seed ^ 0x5DEECE66D) & 0xFFFFFFFFFFFFFFFF, 0, dptr(scratch, torch.float64),
                                              stream_ptr(dev)))
    return train_x, train_y


def main(argv=None):
    """CLI of the reference (signals.py:302-332): ``python -m qbold_vi_b200.signals -f True -b True`` writes
    synthetic_data.npz (x, y) from the ``config`` INI in the working directory (or the packaged defaults)."""
    import argparse
    from .config import load_system_parameters
    parser = argparse.ArgumentParser(description='Generate ASE qBOLD signals')
    parser.add_argument('-f', required=True, help='should the tissue contribution be calculated with the full model')
    parser.add_argument('-b', required=True, help='should the blood contribution be included')
    parser.add_argument('-o', default='synthetic_data', help='output .npz stem')
    args = parser.parse_args(argv)
    if args.f not in ['True', 'False'] or args.b not in ['True', 'False']:
        raise ValueError('Arguments must be a valid boolean')
    params = load_system_parameters()
    # the reference passes misaligned_prob 0.1 here (signals.py:330)
    train_x, train_y = create_synthetic_dataset(params, args.f, args.b, 0.1, False)
    np.savez(args.o, x=train_x.cpu().numpy(), y=train_y.cpu().numpy())


if __name__ == '__main__':
    main()
