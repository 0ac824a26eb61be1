"""The amortization network of the reference (``EncoderTrainer.create_encoder``, model.py:122-223) in PyTorch.

Adjacent to the hot path.  It produces the tensors the fused ELBO kernel consumes -- 5 logit-normal
parameters and ``n_tau`` heteroscedastic sigmas per voxel -- and its 146 176 parameters (optimal.yaml) are what the
per-step NCCL all-reduce carries.  Layout is the reference's channels-last ``[B, X, Y, Z, C]``.  The 3x3x1
convolutions and the forward / input-gradient GEMMs of the per-voxel layers are library kernels (cuDNN / cuBLAS on
tensor cores); hand-written here: the weight / bias gradients of the per-voxel layers (``qbold_dense_wgrad``), the gated
residual mix (``qbold_gate_mix_*``), the whole voxel-wise branch as one tcgen05 kernel for inference
(``voxelwise_fused``) and, opt-in, the single-layer tcgen05 GEMMs (``QBOLD_DENSE_TC=1``).
"""
from __future__ import annotations

import math
import os

import torch

from ._lib import on_device as _on_device
import torch.nn as nn
import torch.nn.functional as F


_TC_STATUS = {}


# The single-layer tcgen05 kernel (qbold_dense_tc) is correct but, at 16 warps per SM, latency-bound at ~75 us per
# 524 288 x 60 layer -- level with cuBLAS (64-104 us), not ahead of it -- so the library GEMMs stay the default for the
# forward / input-gradient passes; QBOLD_DENSE_TC=1 (or this flag) routes them through the kernel.
USE_DENSE_TC = os.environ.get('QBOLD_DENSE_TC', '0') == '1'


def _tc_ok(n_in, n_out, *tensors):
    """Shapes the tcgen05 Dense kernel takes: multiples of 4 up to 64, 16-byte aligned operands."""
    return (n_in % 4 == 0 and n_out % 4 == 0 and 4 <= n_in <= 64 and 4 <= n_out <= 64
            and all(t.data_ptr() % 16 == 0 for t in tensors))


def _small_ok(n_in, n_out, *tensors):
    """The skinny-layer kernels (qbold_dense_small_*): <= 16 outputs, inputs a multiple of 4 up to 64, aligned."""
    return (n_out <= 16 and n_in % 4 == 0 and 4 <= n_in <= 64 and all(t.is_contiguous() and t.data_ptr() % 16 == 0
                                                                      for t in tensors))


def _dense_tc(x, mask, w, bias, n_in, n_out, transpose, relu):
    """y = act((x * [mask > 0]) B^T + bias) through qbold_dense_tc; B = w (transpose=False) or w^T."""
    from . import _lib
    from ._lib import check, dptr, stream_ptr
    lib = _lib.lib()
    dev = x.device
    packed = torch.empty(int(lib.qbold_dense_tc_packed_floats()), dtype=torch.float32, device=dev)
    y = torch.empty((x.shape[0], n_out), dtype=torch.float32, device=dev)
    status = _TC_STATUS.get(dev)
    if status is None:
        status = _TC_STATUS[dev] = torch.zeros(1, dtype=torch.int32, device=dev)
    with _on_device(dev):
        st = stream_ptr(dev)
        check(lib.qbold_dense_tc_pack(dptr(w), dptr(bias, allow_none=True), n_out, n_in, int(transpose), dptr(packed), st))
        check(lib.qbold_dense_tc(dptr(x), dptr(mask, allow_none=True), dptr(packed), n_in, n_out, int(relu), x.shape[0],
                                 dptr(y), dptr(status, torch.int32), st))
    return y


# The TMA-fed pipeline kernel (qbold_dense_tma): 48 us per 524 288 x 60 layer against 67 us for cuBLAS, 67 us against
# 114 us with the in-place accumulation of the backward pass -- the default for the training passes in TF32 mode.
USE_DENSE_TMA = os.environ.get('QBOLD_DENSE_TMA', '1') == '1'


def _tma_ok(n_in, n_out, *tensors):
    return (USE_DENSE_TMA and torch.backends.cuda.matmul.allow_tf32 and n_in % 4 == 0 and n_out % 4 == 0
            and 4 <= n_in <= 64 and 4 <= n_out <= 64
            and all(t is None or (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.data_ptr() % 16 == 0)
                    for t in tensors))


def _dense_tma(x, w, bias=None, addend=None, transpose=False, relu=False):
    """act(x w^T + bias) (transpose=False, w [n_out, n_in]) or x w (+ addend, in place into addend) (transpose=True,
    w [n_in, n_out]) through qbold_dense_tma.  Callers check _tma_ok first."""
    from . import _lib
    from ._lib import check, dptr, stream_ptr
    n = x.shape[0]
    n_in = x.shape[1]
    n_out = w.shape[1] if transpose else w.shape[0]
    dev = x.device
    y = addend if addend is not None else torch.empty((n, n_out), dtype=torch.float32, device=dev)
    status = _TC_STATUS.get(dev)
    if status is None:
        status = _TC_STATUS[dev] = torch.zeros(1, dtype=torch.int32, device=dev)
    with _on_device(dev):
        check(_lib.lib().qbold_dense_tma(dptr(x), dptr(w), dptr(bias, allow_none=True), dptr(addend, allow_none=True), n_in,
                                         n_out, int(transpose), int(relu), n, dptr(y), dptr(status, torch.int32),
                                         stream_ptr(dev)))
    return y


def _linear_relu(x, w, b):
    """relu(x w^T + b) for the block's shared pointwise layer."""
    if _tma_ok(w.shape[1], w.shape[0], x, w, b):
        return _dense_tma(x, w, b, relu=True)
    return torch._addmm_activation(b, x, w.t(), use_gelu=False)


def _linear(x, w, b):
    if _tma_ok(w.shape[1], w.shape[0], x, w, b):
        return _dense_tma(x, w, b)
    return torch.addmm(b, x, w.t())


def _addmm_(acc, g, w):
    """acc += g w (w [n_out, n_in], the stored weight: an input gradient accumulated in place)."""
    if _tma_ok(w.shape[0], w.shape[1], acc, g, w):
        return _dense_tma(g, w, addend=acc, transpose=True)
    return acc.addmm_(g, w)


def _mm(g, w):
    if _tma_ok(w.shape[0], w.shape[1], g, w):
        return _dense_tma(g, w, transpose=True)
    return g @ w


def tensor_core_status(device):
    """Non-zero if a tcgen05 completion barrier of the Dense kernels ever timed out on `device` (checked lazily: the
    training loop never synchronises on it)."""
    s = _TC_STATUS.get(torch.device(device))
    return 0 if s is None else int(s.item())


class _DenseFn(torch.autograd.Function):
    """y = act(x W^T + b) for a per-voxel Dense layer (N ~ 10^5..10^6 rows, <= 64 features).  Forward and input
    gradient run on the tcgen05 kernel (qbold_dense_tc, ReLU and ReLU' fused) when the shape allows, else on library
    GEMMs; the weight / bias gradient -- a reduction over all voxels that cuBLAS serves poorly at this shape -- is
    qbold_dense_wgrad (TF32 mma, HBM-bound, deterministic)."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu=False, input_is_relu=False):
        # input_is_relu: x is itself a ReLU output with this layer as its ONLY consumer; the input gradient is then
        # returned already multiplied by [x > 0] (fused into the skinny kernel's stores), and the producer skips its
        # own ReLU' pass (_BlockFn's premasked_out1)
        ctx.input_is_relu = bool(input_is_relu)
        n_out, n_in = weight.shape
        if not relu and n_out >= 8 and _tma_ok(n_in, n_out, x, weight, bias):      # e.g. the 16-output double head
            y = _dense_tma(x, weight, bias)
        elif not relu and _small_ok(n_in, n_out, x, weight):
            from . import _lib
            from ._lib import check, dptr, stream_ptr
            y = torch.empty((x.shape[0], n_out), dtype=torch.float32, device=x.device)
            with _on_device(x.device):
                check(_lib.lib().qbold_dense_small_forward(dptr(x), dptr(weight), dptr(bias), n_in, n_out, x.shape[0],
                                                           dptr(y), stream_ptr(x.device)))
        elif USE_DENSE_TC and _tc_ok(n_in, n_out, x):
            y = _dense_tc(x, None, weight, bias, n_in, n_out, False, relu)
        elif _tma_ok(n_in, n_out, x, weight, bias):
            y = _dense_tma(x, weight, bias, relu=relu)
        elif relu:                                            # bias + ReLU in the GEMM epilogue (cuBLASLt)
            y = torch._addmm_activation(bias, x, weight.t(), use_gelu=False)
        else:
            y = torch.addmm(bias, x, weight.t())
        if relu:
            ctx.save_for_backward(x, weight, y)
        else:
            ctx.save_for_backward(x, weight)
        ctx.relu = relu
        return y

    @staticmethod
    def backward(ctx, g):
        from . import _lib
        from ._lib import check, dptr, stream_ptr
        if ctx.relu:
            x, weight, y = ctx.saved_tensors
        else:
            (x, weight), y = ctx.saved_tensors, None
        g = g.contiguous()
        n_out, n_in = weight.shape
        dev = x.device
        gx, mask = None, y
        if y is None and n_out >= 8 and _tma_ok(n_out, n_in, g, weight):
            if ctx.needs_input_grad[0]:
                gx = _dense_tma(g, weight, transpose=True)
        elif y is None and _small_ok(n_in, n_out, weight):
            if ctx.needs_input_grad[0]:
                gx = torch.empty((x.shape[0], n_in), dtype=torch.float32, device=dev)
                with _on_device(dev):
                    check(_lib.lib().qbold_dense_small_dgrad_masked(dptr(g), dptr(weight),
                                                                    dptr(x) if ctx.input_is_relu else None, n_in, n_out,
                                                                    x.shape[0], dptr(gx), stream_ptr(dev)))
        elif USE_DENSE_TC and _tc_ok(n_out, n_in, g) and (y is None or y.data_ptr() % 16 == 0):
            if ctx.needs_input_grad[0]:
                gx = _dense_tc(g, y, weight, None, n_out, n_in, True, False)          # (g * relu') W, relu' fused
        else:
            if y is not None and ctx.needs_input_grad[0]:     # materialise g * relu' once for both uses
                g, mask = torch.ops.aten.threshold_backward(g, y, 0.0), None
            if ctx.needs_input_grad[0]:
                gx = _mm(g, weight)
        lib = _lib.lib()
        if not torch.backends.cuda.matmul.allow_tf32:         # strict float32 requested: qbold_dense_wgrad is TF32 mma
            gm = g if mask is None else torch.ops.aten.threshold_backward(g, mask, 0.0)
            return gx, gm.t() @ x, gm.sum(0), None, None
        if mask is None and _tma_ok(n_in, n_out, g, x):
            dw, db = _wgrad(g, x)
            return gx, dw, db, None, None
        ws = _workspace('dense_wgrad', dev)
        dw = torch.empty_like(weight)
        db = torch.empty(n_out, dtype=torch.float32, device=dev)
        with _on_device(dev):
            check(lib.qbold_dense_wgrad(dptr(g), dptr(mask, allow_none=True), n_out, dptr(x), n_in, x.shape[0], dptr(dw),
                                        dptr(db), 0, dptr(ws), stream_ptr(dev)))
        return gx, dw, db, None, None


class _GateMixFn(torch.autograd.Function):
    """skip * (1 - g) + r * g with g = sigmoid(z + offset) (model.py:160-172): one streaming kernel each way."""

    @staticmethod
    def forward(ctx, skip, r, z, offset):
        from . import _lib
        from ._lib import check, dptr, stream_ptr
        skip, r, z = skip.contiguous(), r.contiguous(), z.contiguous()
        c, zc = r.shape[-1], z.shape[-1]
        n = r.numel() // c
        out = torch.empty_like(r)
        with _on_device(r.device):
            check(_lib.lib().qbold_gate_mix_forward(dptr(skip), dptr(r), dptr(z), float(offset), n, c, zc, dptr(out),
                                                    stream_ptr(r.device)))
        ctx.save_for_backward(skip, r, z)
        ctx.offset = float(offset)
        return out

    @staticmethod
    def backward(ctx, go):
        from . import _lib
        from ._lib import check, dptr, stream_ptr
        skip, r, z = ctx.saved_tensors
        go = go.contiguous()
        c, zc = r.shape[-1], z.shape[-1]
        n = r.numel() // c
        d_skip, d_r, d_z = torch.empty_like(skip), torch.empty_like(r), torch.empty_like(z)
        with _on_device(r.device):
            check(_lib.lib().qbold_gate_mix_backward(dptr(go), dptr(skip), dptr(r), dptr(z), ctx.offset, n, c, zc,
                                                     dptr(d_skip), dptr(d_r), dptr(d_z), stream_ptr(r.device)))
        return d_skip, d_r, d_z, None


def gate_mix(skip, r, z, offset):
    if r.is_cuda and r.dtype == torch.float32 and skip.dtype == torch.float32 and z.dtype == torch.float32:
        return _GateMixFn.apply(skip, r, z, offset)
    g = torch.sigmoid(z + offset)
    return skip * (1.0 - g) + r * g


def dense(layer, x, relu=False):
    """nn.Linear on the last axis (optionally followed by ReLU); routes through _DenseFn when training on CUDA in
    TF32 mode with a supported shape (weights contiguous float32, <= 64 outputs, <= 63 inputs), else F.linear."""
    w, b = layer.weight, layer.bias
    if (x.is_cuda and torch.is_grad_enabled() and w.requires_grad and b is not None and x.dtype == torch.float32
            and torch.backends.cuda.matmul.allow_tf32 and w.shape[0] <= 64 and w.shape[1] <= 63
            and not torch.is_autocast_enabled()):
        lead = x.shape[:-1]
        return _DenseFn.apply(x.reshape(-1, x.shape[-1]).contiguous(), w, b, relu).reshape(lead + (w.shape[0],))
    return F.relu(layer(x)) if relu else layer(x)


def _he_normal_(w, fan_in):
    # keras HeNormal: truncated normal, stddev = sqrt(2 / fan_in) / 0.87962566
    std = math.sqrt(2.0 / fan_in) / 0.87962566103423978
    nn.init.trunc_normal_(w, mean=0.0, std=std, a=-2 * std, b=2 * std)


class _Conv331(nn.Module):
    """keras Conv3D(kernel_size=(3,3,1), padding='same') on a channels-last volume.  The reference layout
    [B, X, Y, Z, C] IS torch's channels_last_3d layout of a [B, C, X, Y, Z] tensor, so the permutes below are views
    (no transposes hit HBM) and cuDNN runs its NDHWC kernels; the kernel never mixes z slices, so z-slabs need no
    halo when volumes are sharded."""

    def __init__(self, c_in, c_out, std):
        super().__init__()
        self.conv = nn.Conv3d(c_in, c_out, (3, 3, 1), padding=(1, 1, 0))
        nn.init.normal_(self.conv.weight, std=std)
        nn.init.zeros_(self.conv.bias)
        self.conv.weight.data = self.conv.weight.data.contiguous(memory_format=torch.channels_last_3d)

    def forward(self, x):                                   # [B, X, Y, Z, C], contiguous
        y = self.conv(x.permute(0, 4, 1, 2, 3))             # view: [B, C, X, Y, Z] in channels_last_3d strides
        return y.permute(0, 2, 3, 4, 1)


class _Block(nn.Module):
    """create_block (model.py:142-174): stream-1 1x1x1 conv shared with the skip of stream 2, plus a gated
    residual pair of 3x3x1 convs."""

    def __init__(self, units, act, resid_std, channelwise_gating, gate_offset):
        super().__init__()
        self.act = act
        self.gate_offset = gate_offset
        self.pointwise = nn.Linear(units, units)
        _he_normal_(self.pointwise.weight, units)
        nn.init.zeros_(self.pointwise.bias)
        self.conv_a = _Conv331(units, units, resid_std)
        self.conv_b = _Conv331(units, units, resid_std)
        self.gate = nn.Linear(units, units if channelwise_gating else 1)
        nn.init.normal_(self.gate.weight, std=resid_std)
        nn.init.zeros_(self.gate.bias)

    def forward(self, net1, net2):
        fuse = self.act is F.relu
        out1 = dense(self.pointwise, net1, True) if fuse else self.act(dense(self.pointwise, net1))
        skip = dense(self.pointwise, net2, True) if fuse else self.act(dense(self.pointwise, net2))
        r = self.conv_b(self.act(self.conv_a(self.act(net2))))
        return out1, gate_mix(skip, r, dense(self.gate, r), self.gate_offset)


# ---------------------------------------------------------------------------------------------------------------------
# Fused training path of the gated residual block (create_block, model.py:142-174).
#
# Internally the encoder keeps its activations as [B, Z, X, Y, C] ("z-outer"): the reference's 3x3x1 convolutions never
# mix z slices, so in this layout they are plain 3x3 2-D convolutions over a batch of B*Z channels-last images -- the
# shape cuDNN's 2-D NHWC kernels are tuned for (76 us against 152 us for the 3-D NDHWC form at 2 x 64^3 x 60, TF32) --
# and the per-voxel Dense layers do not care.  Only the 11-channel input and the 5 / 5 / 11-channel outputs are
# transposed.  One autograd node per block runs: the shared pointwise Dense (+ReLU, cuBLASLt epilogue), conv + bias +
# ReLU fused in cuDNN, the second conv WITHOUT its bias (its bias is folded into the gate Dense and the mix kernel:
# cuDNN would add it in a separate 86 us pass), the gate Dense and the mix kernel; the backward uses the fused
# streaming kernels of csrc/encoder_block_kernels.cuh (ReLU' + bias gradient in one pass, ReLU' of the skip branch inside the
# mix backward), in-place accumulating GEMMs (beta = 1) instead of separate gradient additions, and qbold_dense_wgrad.
_FAST_BLOCK = os.environ.get('QBOLD_FAST_BLOCK', '1') == '1'


_WORKSPACES = {}


def _workspace(kind, dev):
    """Device scratch of one of the reduction kernels, allocated once per (kind, device, stream): launches on one stream
    are ordered, so consecutive calls can share it; a different stream (a concurrent backward pass) gets its own."""
    from . import _lib
    key = (kind, dev.index, _lib.stream_ptr(dev).value)
    ws = _WORKSPACES.get(key)
    if ws is None:
        lib = _lib.lib()
        n = {'colsum': lib.qbold_colsum_workspace_floats, 'dense_wgrad': lib.qbold_dense_wgrad_workspace_floats,
             'dense_wgrad_tma': lib.qbold_dense_wgrad_tma_workspace_floats,
             'conv_wgrad': lib.qbold_conv_wgrad_workspace_floats}[kind]()
        ws = _WORKSPACES[key] = torch.empty(int(n), dtype=torch.float32, device=dev)
    return ws


def _ws(dev):
    return _workspace('colsum', dev)


def _relu_bwd(g, y, addend=None, colsum=None):
    """g * [y > 0] (+ addend), optionally with the column sums of the result (a bias gradient) from the same pass."""
    from . import _lib
    from ._lib import check, dptr, stream_ptr
    n, c = g.shape
    out = torch.empty_like(g)
    ws = _ws(g.device) if colsum is not None else None
    with _on_device(g.device):
        check(_lib.lib().qbold_relu_bwd_colsum(dptr(g), dptr(y), dptr(addend, allow_none=True), n, c, dptr(out),
                                               dptr(colsum, allow_none=True), 0, dptr(ws, allow_none=True),
                                               stream_ptr(g.device)))
    return out


def _colsum(g):
    from . import _lib
    from ._lib import check, dptr, stream_ptr
    n, c = g.shape
    out = torch.empty(c, dtype=torch.float32, device=g.device)
    ws = _ws(g.device)
    with _on_device(g.device):
        check(_lib.lib().qbold_relu_bwd_colsum(dptr(g), None, None, n, c, None, dptr(out), 0, dptr(ws), stream_ptr(g.device)))
    return out


def _wgrad(g, x, accumulate_into=None):
    """dW [n_out, n_in] = g^T x and db = column sums of g (qbold_dense_wgrad); optionally added to an earlier pair."""
    from . import _lib
    from ._lib import check, dptr, stream_ptr
    lib = _lib.lib()
    n_out, n_in = g.shape[1], x.shape[1]
    dev = g.device
    if not torch.backends.cuda.matmul.allow_tf32 or n_in > 63:      # strict float32 requested: the kernel is TF32 mma
        dw, db = g.t() @ x, g.sum(0)
        if accumulate_into is not None:
            accumulate_into[0].add_(dw)
            accumulate_into[1].add_(db)
            return accumulate_into
        return dw, db
    if accumulate_into is None:
        dw = torch.empty((n_out, n_in), dtype=torch.float32, device=dev)
        db = torch.empty(n_out, dtype=torch.float32, device=dev)
    else:
        dw, db = accumulate_into
    if _tma_ok(n_in, n_out, g, x):                          # TMA-fed tcgen05 kernel (MN-major operands)
        ws = _workspace('dense_wgrad_tma', dev)
        status = _TC_STATUS.get(dev)
        if status is None:
            status = _TC_STATUS[dev] = torch.zeros(1, dtype=torch.int32, device=dev)
        with _on_device(dev):
            check(lib.qbold_dense_wgrad_tma(dptr(g), n_out, dptr(x), n_in, x.shape[0], dptr(dw), dptr(db),
                                            0 if accumulate_into is None else 1, dptr(ws), dptr(status, torch.int32),
                                            stream_ptr(dev)))
        return dw, db
    ws = _workspace('dense_wgrad', dev)
    with _on_device(dev):
        check(lib.qbold_dense_wgrad(dptr(g), None, n_out, dptr(x), n_in, x.shape[0], dptr(dw), dptr(db),
                                    0 if accumulate_into is None else 1, dptr(ws), stream_ptr(dev)))
    return dw, db


def _as_images(flat, dims):
    """[N, C] z-outer activations as the [B*Z, C, X, Y] channels-last view cuDNN's 2-D kernels take (no copy)."""
    bz, x, y = dims
    return flat.view(bz, x, y, flat.shape[1]).permute(0, 3, 1, 2)


def _as_flat(img):
    """Back to [N, C]; a copy only if the library returned a non-channels-last tensor."""
    return img.permute(0, 2, 3, 1).reshape(-1, img.shape[1])


_CONV_ARGS = ((1, 1), (1, 1), (1, 1))                 # stride, padding, dilation
# weight gradient of the 3x3 convolutions on tcgen05 (qbold_conv_wgrad) instead of cuDNN's wgrad kernel; TF32, so only
# when TF32 convolutions are allowed
_CONV_WGRAD_TC = os.environ.get('QBOLD_CONV_WGRAD_TC', '1') == '1'


def _conv_backward(g_flat, x_flat, w2, dims):
    """(input gradient [N, C_in] flat, weight gradient [C_out, C_in, 3, 3]) of a 3x3 'same' convolution on z-outer
    activations: the input gradient is cuDNN's, the weight gradient the tcgen05 kernel (or cuDNN's in strict float32)."""
    from . import _lib
    from ._lib import check, dptr, stream_ptr
    bz, nx, ny = dims
    c_out, c_in = g_flat.shape[1], x_flat.shape[1]
    use_tc = (_CONV_WGRAD_TC and torch.backends.cudnn.allow_tf32 and c_out % 4 == 0 and c_in % 4 == 0 and c_out <= 64
              and c_in <= 64 and g_flat.is_contiguous() and x_flat.is_contiguous())
    mask = (True, not use_tc, False)
    d_in, dw, _ = torch.ops.aten.convolution_backward(_as_images(g_flat, dims), _as_images(x_flat, dims), w2, None,
                                                      (1, 1), (1, 1), (1, 1), False, (0, 0), 1, mask)
    if use_tc:
        lib = _lib.lib()
        dev = g_flat.device
        dw = torch.empty((c_out, c_in, 3, 3), dtype=torch.float32, device=dev)
        ws = _workspace('conv_wgrad', dev)
        status = _TC_STATUS.get(dev)
        if status is None:
            status = _TC_STATUS[dev] = torch.zeros(1, dtype=torch.int32, device=dev)
        with _on_device(dev):
            check(lib.qbold_conv_wgrad(dptr(g_flat), c_out, dptr(x_flat), c_in, bz, nx, ny, dptr(dw), 0, dptr(ws),
                                       dptr(status, torch.int32), stream_ptr(dev)))
    return _as_flat(d_in), dw


class _BlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net1, net2, a0, w_p, b_p, w_a, b_a, w_b, b_b, w_g, b_g, dims, offset, same_input, a0_is_net2,
                want_relu_out, premasked_out1=False):
        from . import _lib
        from ._lib import check, dptr, stream_ptr
        n, c = net2.shape
        dev = net2.device
        skip = _linear_relu(net2, w_p, b_p)
        out1 = skip if same_input else _linear_relu(net1, w_p, b_p)
        wa2 = w_a.squeeze(-1).contiguous(memory_format=torch.channels_last)
        wb2 = w_b.squeeze(-1).contiguous(memory_format=torch.channels_last)
        c1 = torch.cudnn_convolution_relu(_as_images(a0, dims), wa2, b_a, *_CONV_ARGS, 1)          # conv + bias + ReLU
        r0 = F.conv2d(c1, wb2, None, padding=1)                                                    # bias folded below
        c1f, r0f = _as_flat(c1), _as_flat(r0)
        z = _linear(r0f, w_g, torch.addmv(b_g, w_g, b_b))                                          # W_g (r0 + b_b) + b_g
        out2 = torch.empty_like(r0f)
        out2_relu = torch.empty_like(r0f) if want_relu_out else None
        with _on_device(dev):
            check(_lib.lib().qbold_block_mix_forward(dptr(skip), dptr(r0f), dptr(b_b.contiguous()), dptr(z), float(offset), n,
                                                     c, dptr(out2), dptr(out2_relu, allow_none=True), stream_ptr(dev)))
        ctx.save_for_backward(net1, net2, a0, out1, skip, c1f, r0f, z, w_p, w_a, w_b, w_g, b_b)
        ctx.dims, ctx.offset, ctx.same_input, ctx.a0_is_net2 = dims, float(offset), same_input, a0_is_net2
        ctx.wa2, ctx.wb2 = wa2, wb2                      # channels-last copies of the 3x3 weights, reused by backward
        ctx.premasked_out1 = bool(premasked_out1)       # the gradient of out1 arrives already times [out1 > 0]
        if want_relu_out:
            ctx.mark_non_differentiable(out2_relu)
            return out1, out2, out2_relu
        return out1, out2, None

    @staticmethod
    def backward(ctx, d_out1, d_out2, _unused):
        from . import _lib
        from ._lib import check, dptr, stream_ptr
        net1, net2, a0, out1, skip, c1f, r0f, z, w_p, w_a, w_b, w_g, b_b = ctx.saved_tensors
        n, c = net2.shape
        dev = net2.device
        dims = ctx.dims
        d_out2 = d_out2.contiguous()
        d_skip, d_r, d_z = torch.empty_like(skip), torch.empty_like(skip), torch.empty_like(skip)
        # block 0: out1 IS skip, so the stream-1 gradient joins the skip gradient inside the mix kernel, before its ReLU'
        # mask -- d_skip then already is g_p = [skip > 0] * (d_out2 (1 - g) + d_out1)
        d_out1c = d_out1.contiguous() if ctx.same_input else None
        with _on_device(dev):
            check(_lib.lib().qbold_block_mix_backward_add(dptr(d_out2), dptr(skip), dptr(r0f), dptr(b_b.contiguous()),
                                                          dptr(z), ctx.offset, n, c, 1, dptr(d_out1c, allow_none=True),
                                                          dptr(d_skip), dptr(d_r), dptr(d_z), stream_ptr(dev)))
        # gate Dense: z = W_g (r0 + b_b) + b_g
        dw_g, db_g = _wgrad(d_z, r0f)
        dw_g = torch.addr(dw_g, db_g, b_b)                                   # + colsum(d_z) (x) b_b
        _addmm_(d_r, d_z, w_g)                                               # total gradient of r = r0 + b_b, in place
        db_b = _colsum(d_r)
        # second convolution (no bias of its own): input gradient + weight gradient
        wa2, wb2 = ctx.wa2, ctx.wb2
        d_c1, dw_b = _conv_backward(d_r, c1f, wb2, dims)
        db_a = torch.empty(c, dtype=torch.float32, device=dev)
        d_c1m = _relu_bwd(d_c1, c1f, colsum=db_a)                            # ReLU' and the bias gradient in one pass
        d_net2, dw_a = _conv_backward(d_c1m, a0, wa2, dims)
        if not ctx.a0_is_net2:                                               # a0 = relu(net2): apply its derivative
            d_net2 = _relu_bwd(d_net2, net2)
        elif not d_net2.is_contiguous():
            d_net2 = d_net2.contiguous()
        # shared pointwise Dense: skip = relu(W_p net2 + b_p) (d_skip is already masked), out1 = relu(W_p net1 + b_p)
        if ctx.same_input:
            g_p = d_skip                                                     # both uses of the same activation
            dw_p, db_p = _wgrad(g_p, net2)
            _addmm_(d_net2, g_p, w_p)
            d_net1 = None
        else:
            g_1 = d_out1.contiguous() if ctx.premasked_out1 else _relu_bwd(d_out1.contiguous(), out1)
            dw_p, db_p = _wgrad(d_skip, net2)
            _wgrad(g_1, net1, accumulate_into=(dw_p, db_p))
            _addmm_(d_net2, d_skip, w_p)
            d_net1 = _mm(g_1, w_p)
        return (d_net1, d_net2, None, dw_p, db_p, dw_a.unsqueeze(-1), db_a, dw_b.unsqueeze(-1), db_b, dw_g, db_g,
                None, None, None, None, None, None)


def _fast_block_ok(blk, net2):
    u = blk.pointwise.out_features
    return (_FAST_BLOCK and net2.is_cuda and net2.dtype == torch.float32 and blk.act is F.relu and u % 4 == 0 and u <= 64
            and blk.gate.out_features == u and torch.backends.cudnn.is_available() and not torch.is_autocast_enabled() and torch.is_grad_enabled()
            and hasattr(torch, 'cudnn_convolution_relu'))


class Encoder(nn.Module):
    """outer_model of create_encoder: data [B,X,Y,Z,n_tau] -> (q_voxelwise [...,5], q_spatial [...,5], sigma [...,n_tau])."""

    def __init__(self, no_units=60, no_intermediate_layers=2, activation='relu', initial_im_sigma=0.05,
                 multi_image_normalisation=False, channelwise_gating=True, gate_offset=-3.0, resid_init_std=0.05,
                 no_ip_images=11, se_idx=2, use_mvg=True, infer_inv_gamma=False):
        super().__init__()
        # infer_inv_gamma (model.py:201-205): a learned 4-vector exp(v), v initialised to log([20, 2.5, 20, 2.5]), is
        # broadcast onto output 0 as 4 extra channels (alpha_oef, beta_oef, alpha_dbv, beta_dbv)
        self.hyper_prior = nn.Parameter(torch.log(torch.tensor([20.0, 2.5, 20.0, 2.5]))) if infer_inv_gamma else None
        self.act = {'relu': F.relu, 'gelu': F.gelu, 'tanh': torch.tanh}[activation]
        self.se_idx = se_idx
        self.multi_image_normalisation = multi_image_normalisation
        self.first = nn.Linear(no_ip_images, no_units)
        _he_normal_(self.first.weight, no_ip_images)
        nn.init.zeros_(self.first.bias)
        self.blocks = nn.ModuleList([_Block(no_units, self.act, resid_init_std, channelwise_gating, gate_offset)
                                     for _ in range(max(1, no_intermediate_layers))])
        self.final = nn.Linear(no_units, 5 if use_mvg else 4)
        _he_normal_(self.final.weight, no_units)
        nn.init.zeros_(self.final.bias)
        self.im_sigma = nn.Linear(no_units, no_ip_images)
        nn.init.normal_(self.im_sigma.weight, std=resid_init_std)
        nn.init.constant_(self.im_sigma.bias, math.log(initial_im_sigma))

    def normalise_data(self, data):
        """model.py:97-113: clip, divide by the tau=0 image (or the 3-image mean), log."""
        d = torch.clamp(data, 1e-2, 1e8)
        se = self.se_idx
        ref = d[..., se - 1:se + 2] if self.multi_image_normalisation else d[..., se:se + 1]
        return torch.log(d / ref.mean(-1, keepdim=True))

    def supports_voxelwise_fused(self):
        """Limits of the tensor-core kernel: ReLU, <= 64 units, <= 32 input images, 1..6 blocks, <= 16 outputs."""
        return (self.act is F.relu and self.first.out_features <= 64 and self.first.in_features <= 32 and
                1 <= len(self.blocks) <= 6 and self.final.out_features <= 16)

    @torch.no_grad()
    def voxelwise_fused(self, data):
        """q_voxelwise (output 0 of forward) from ONE tcgen05 kernel (qbold_encoder_mlp_forward): normalise_data, the
        first Dense, the blocks' stream-1 Dense layers and the final Dense, TF32 tensor cores with fp32 accumulation
        in TMEM, inference only.  Needs ReLU, no_units <= 64, n_tau <= 32."""
        import ctypes as C
        from . import _lib
        from ._lib import check, dptr, stream_ptr
        if self.act is not F.relu:
            raise _lib.QboldError('voxelwise_fused supports the ReLU encoder (optimal.yaml) only')
        lead = tuple(data.shape[:-1])
        x = data.reshape(-1, data.shape[-1]).float().contiguous()
        n, n_in = x.shape
        n_mid, n_out, hidden = len(self.blocks), self.final.out_features, self.first.out_features
        lib = _lib.lib()
        n_blob = lib.qbold_encoder_mlp_blob_floats(n_mid)
        if n_blob < 0:
            raise _lib.QboldError('voxelwise_fused supports 1..6 blocks')
        dev = x.device
        blob = torch.empty(n_blob, dtype=torch.float32, device=dev)
        q = torch.empty((n, n_out), dtype=torch.float32, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        ws = [b.pointwise.weight.detach().float().contiguous() for b in self.blocks]
        bs = [b.pointwise.bias.detach().float().contiguous() for b in self.blocks]
        w_arr = (C.c_void_p * n_mid)(*[w.data_ptr() for w in ws])
        b_arr = (C.c_void_p * n_mid)(*[b.data_ptr() for b in bs])
        w_in, b_in = self.first.weight.detach().float().contiguous(), self.first.bias.detach().float().contiguous()
        w_out, b_out = self.final.weight.detach().float().contiguous(), self.final.bias.detach().float().contiguous()
        with _on_device(dev):
            st = stream_ptr(dev)
            check(lib.qbold_encoder_mlp_pack(dptr(w_in), dptr(b_in), w_arr, b_arr, dptr(w_out), dptr(b_out), n_in, hidden,
                                             n_mid, n_out, dptr(blob), st))
            check(lib.qbold_encoder_mlp_forward(dptr(x), dptr(blob), n_in, n_mid, n_out, self.se_idx,
                                                int(self.multi_image_normalisation), n, dptr(q),
                                                dptr(status, torch.int32), st))
        if int(status.item()) != 0:
            raise _lib.QboldError('k_encoder_mlp: a tensor-core completion barrier timed out')
        return q.reshape(lead + (n_out,))

    def forward_voxelwise(self, data):
        """Output 0 alone through the torch layers (differentiable): the only output the pre-training loss uses
        (train.py:402-412 compiles the model with a loss on the first output), so the 3x3x1 stream is skipped."""
        fuse = self.act is F.relu
        x = self.normalise_data(data)
        h = dense(self.first, x, True) if fuse else self.act(dense(self.first, x))
        for blk in self.blocks:
            h = dense(blk.pointwise, h, True) if fuse else self.act(dense(blk.pointwise, h))
        return self._with_hyper_prior(dense(self.final, h))

    def _with_hyper_prior(self, out):
        if self.hyper_prior is None:
            return out
        return torch.cat([out, torch.exp(self.hyper_prior).expand(out.shape[:-1] + (4,))], -1)     # model.py:205

    def forward(self, data):
        if data.dim() == 5 and _fast_block_ok(self.blocks[0], data):
            return self._forward_fused_blocks(data)
        x = self.normalise_data(data)
        h = dense(self.first, x, True) if self.act is F.relu else self.act(dense(self.first, x))
        net1 = net2 = h
        for blk in self.blocks:
            net1, net2 = blk(net1, net2)
        return (self._with_hyper_prior(dense(self.final, net1)), dense(self.final, net2),
                torch.exp(dense(self.im_sigma, net2)))

    def _forward_fused_blocks(self, data):
        """Training path on CUDA (TF32, ReLU, channel-wise gating): z-outer layout + one fused autograd node per block."""
        b, nx, ny, nz, n_tau = data.shape
        dims = (b * nz, nx, ny)
        tp = (n_tau + 3) & ~3
        if (not data.requires_grad and n_tau <= 64 and data.numel() > 0
                and _tma_ok(tp, self.first.out_features, self.first.weight, self.first.bias)):
            # normalise_data + the move to z-outer rows in ONE pass (rows padded to a multiple of 4 images: 16-byte
            # aligned operands), then the first Dense layer on the TMA-fed tensor-core kernel with zero-padded weights
            from . import _lib
            from ._lib import check, dptr, stream_ptr
            src = data.contiguous()
            xt = torch.empty((b * nz * nx * ny, tp), dtype=torch.float32, device=data.device)
            with _on_device(data.device):
                check(_lib.lib().qbold_normalise_zouter(dptr(src), b, nx, ny, nz, n_tau, self.se_idx,
                                                        int(self.multi_image_normalisation), dptr(xt),
                                                        stream_ptr(data.device)))
            h = _DenseFn.apply(xt, F.pad(self.first.weight, (0, tp - n_tau)), self.first.bias, True)
        else:
            x = self.normalise_data(data)
            xt = x.permute(0, 3, 1, 2, 4).reshape(b * nz * nx * ny, n_tau)            # [B,Z,X,Y,n_tau]: 44 B/voxel copy
            h = dense(self.first, xt, True)
        net1 = net2 = a0 = h
        # the stream-1 head is the only reader of the last block's out1 (a ReLU output): its skinny input-gradient kernel
        # applies [out1 > 0] itself, so the block skips a ReLU' pass (needs the float32 skinny path: <= 16 outputs)
        head_masks = _small_ok(self.final.in_features, self.final.out_features, self.final.weight)
        for i, blk in enumerate(self.blocks):
            last = i + 1 == len(self.blocks)
            net1, net2, nxt = _BlockFn.apply(net1, net2, a0, blk.pointwise.weight, blk.pointwise.bias, blk.conv_a.conv.weight,
                                             blk.conv_a.conv.bias, blk.conv_b.conv.weight, blk.conv_b.conv.bias,
                                             blk.gate.weight, blk.gate.bias, dims, blk.gate_offset, i == 0, i == 0,
                                             not last, last and i > 0 and head_masks)
            a0 = nxt

        def back(t):                                                                  # [B,Z,X,Y,c] -> [B,X,Y,Z,c]
            return t.view(b, nz, nx, ny, t.shape[-1]).permute(0, 2, 3, 1, 4).contiguous()

        if head_masks and len(self.blocks) > 1:
            q1 = _DenseFn.apply(net1, self.final.weight, self.final.bias, False, True)
        else:
            q1 = dense(self.final, net1)
        # the two heads that read net2 (posterior parameters and sigmas) as ONE skinny Dense: one pass over net2 forward,
        # one input-gradient pass backward
        n_q = self.final.out_features
        w_cat = torch.cat([self.final.weight, self.im_sigma.weight], 0)
        b_cat = torch.cat([self.final.bias, self.im_sigma.bias], 0)
        if w_cat.shape[0] <= 16 and net2.is_cuda:
            both = _DenseFn.apply(net2, w_cat, b_cat, False)
            q2, sg = both[:, :n_q], torch.exp(both[:, n_q:])
        else:
            q2, sg = dense(self.final, net2), torch.exp(dense(self.im_sigma, net2))
        return self._with_hyper_prior(back(q1)), back(q2), back(sg)


def create_encoder_from_args(args, no_ip_images=11, se_idx=2):
    """train.py:430-451 with an argparse/yaml namespace (config.load_arguments)."""
    if getattr(args, 'use_layer_norm', False) or float(getattr(args, 'dropout_rate', 0.0) or 0.0) > 0.0:
        raise NotImplementedError('use_layer_norm / dropout_rate > 0 (model.py:136-147) are not provided by this encoder; '
                                  'optimal.yaml uses neither')
    return Encoder(no_units=max(1, args.no_units), no_intermediate_layers=max(1, args.no_intermediate_layers),
                   activation=args.activation, initial_im_sigma=args.im_loss_sigma,
                   multi_image_normalisation=args.multi_image_normalisation,
                   channelwise_gating=args.channelwise_gating, gate_offset=args.gate_offset,
                   resid_init_std=args.resid_init_std, no_ip_images=no_ip_images, se_idx=se_idx,
                   use_mvg=args.use_mvg, infer_inv_gamma=bool(getattr(args, 'infer_inv_gamma', False)))
