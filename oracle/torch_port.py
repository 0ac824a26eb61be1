"""Op-for-op PyTorch-CPU port of the reference's forward model + autodiff (signals.py:98-114,
152-193, 233-247), used as the multi-threaded CPU baseline and as a second check of the NumPy
oracle.  TEST / BENCH INFRASTRUCTURE ONLY.

Same tensor program TensorFlow runs on CPU: the [N, n_tau, 129] Bessel argument is materialised,
``bessel_j0`` has the registered gradient ``-bessel_j1`` (TF math_grad), everything float32.
torch.special.bessel_j0/j1 (float32 CPU) differ from the Cephes single-precision kernels TF uses by
<= 4.2e-7 abs -- fine for a timing baseline and a 1e-5 cross-check.
"""
import math
import time

import torch


class _BesselJ0(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.special.bessel_j0(x)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return -torch.special.bessel_j1(x) * dy


def forward(ph, oef_dbv):
    """ph: oracle.qbold_oracle.Physics; oef_dbv: float32 torch tensor [N,2] (may require grad)."""
    f32 = torch.float32
    oef, dbv = oef_dbv[:, 0:1], oef_dbv[:, 1:2]
    taus = torch.as_tensor(ph.taus, dtype=f32)
    K = (4.0 / 3.0) * math.pi * ph.gamma * ph.b0 * ph.dchi * ph.hct
    dw = K * oef
    a, b = torch.tensor(1e-5, dtype=f32), torch.tensor(1.0, dtype=f32)
    delta = (b - a) / 128.0
    u = torch.cat([a.reshape(1), a + delta * torch.arange(1, 128, dtype=f32), b.reshape(1)])
    arg = 1.5 * (taus[None, :] * dw).unsqueeze(-1) * u                               # [N, n_tau, 129]
    y = (2 + u) * torch.sqrt(1 - u) * (1.0 - _BesselJ0.apply(arg)) / (3.0 * u * u)
    h = (u[2] - u[0]) / 2.0
    integ = ((y[..., 0:-2:2] + y[..., 2::2] + 4.0 * y[..., 1:-1:2]) * (h / 3.0)).sum(-1)
    tissue = torch.exp(-dbv * integ) * math.exp(-ph.te * ph.r2t)
    m_bld = 1 - (2 - math.exp(-(ph.tr - ph.ti) / ph.t1b)) * math.exp(-ph.ti / ph.t1b)
    bw = m_bld * 0.775 * dbv
    td = ((2.6 ** 2.0) / 2.0) * 1e-3
    g0 = (4 / 45) * ph.hct * (1 - ph.hct) * (4.0 * math.pi * ph.b0 * ph.dchi * oef) ** 2
    B = (ph.te / td) + math.sqrt(0.25 + ph.te / td) + 1.5 - 2.0 * torch.sqrt(0.25 + (ph.te + taus) / td) \
        - 2.0 * torch.sqrt(0.25 + (ph.te - taus) / td)
    blood = math.exp(-(1.0 / 0.189) * ph.te) * torch.exp(-(0.5 * ph.gamma ** 2 * g0 * td ** 2) * B)
    return (1 - bw) * tissue + bw * blood


def forward_backward(ph, oef_dbv, g_signal, chunk=8192):
    """Returns (signal, grad) as tensors; chunked like the reference chunks generation (signals.py:281-285)."""
    sigs, grads = [], []
    for i in range(0, oef_dbv.shape[0], chunk):
        x = oef_dbv[i:i + chunk].clone().requires_grad_(True)
        s = forward(ph, x)
        (g,) = torch.autograd.grad((s * g_signal[i:i + chunk]).sum(), x)
        sigs.append(s.detach())
        grads.append(g)
    return torch.cat(sigs), torch.cat(grads)


def timed(ph, oef_dbv, g_signal, threads):
    torch.set_num_threads(threads)
    t0 = time.perf_counter()
    forward_backward(ph, oef_dbv, g_signal)
    return time.perf_counter() - t0
