import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_collection_modifyitems(config, items):
    # GPU tests never run without a device: fail loudly instead of silently passing on a fallback
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason='no CUDA device in this container')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def rel_max(a, b):
    """max |a-b| / max |b| : relative error at the scale of the array."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def rel_elem(a, b, floor=0.0):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / (np.abs(b) + floor)))


@pytest.fixture(scope='session')
def cfg_noise_off():
    from oracle import qbold_oracle as o
    c = o.default_config()
    c['simulate_noise'] = 'False'
    return c


@pytest.fixture(scope='session')
def physics(cfg_noise_off):
    from oracle import qbold_oracle as o
    return o.parse_params(cfg_noise_off)


def load_reference_encoder_weights(enc, fix):
    """Install the kernels / biases of the reference-source encoder fixture (keras layout [kx,ky,kz,C_in,C_out], creation
    order: first | (pointwise, conv_a, conv_b, gate) per block | final | im_sigma) into a qbold_vi_b200 Encoder; returns
    the parameters in the same order as the fixture's gradients [(weight, bias, kind), ...]."""
    import torch
    mods = [(enc.first, 'dense')]
    for blk in enc.blocks:
        mods += [(blk.pointwise, 'dense'), (blk.conv_a.conv, 'conv'), (blk.conv_b.conv, 'conv'), (blk.gate, 'dense')]
    mods += [(enc.final, 'dense'), (enc.im_sigma, 'dense')]
    assert len(mods) == int(fix['n_layers'])
    out = []
    with torch.no_grad():
        for i, (m, kind) in enumerate(mods):
            k = torch.as_tensor(fix['kernel%d' % i])
            w = k.permute(4, 3, 0, 1, 2) if kind == 'conv' else k[0, 0, 0].t()
            m.weight.copy_(w.to(m.weight.device))
            m.bias.copy_(torch.as_tensor(fix['bias%d' % i]).to(m.bias.device))
            out.append((m.weight, m.bias, kind))
    return out


def reference_encoder_grad(fix, i, kind):
    import torch
    k = torch.as_tensor(fix['grad_kernel%d' % i])
    return (k.permute(4, 3, 0, 1, 2) if kind == 'conv' else k[0, 0, 0].t()), torch.as_tensor(fix['grad_bias%d' % i])
