"""GPU suite: the CUDA path (through the ctypes C-ABI) against the oracle and the golden fixtures.

Tolerances are the ones BASELINE.json states: signals within 1e-5 relative (FP32), ELBO and its
gradients within 1e-4 relative.  /root/reference is never touched here.
"""
import numpy as np
import pytest
import torch

from conftest import golden, rel_elem, rel_max
from oracle import philox
from oracle import qbold_oracle as o

pytestmark = pytest.mark.gpu

SIG_TOL = 1e-5
GRAD_TOL = 1e-4
ELEM_TOL = 1e-4            # element-wise, against max(|ref|, 1e-3 * max|ref|); see _elem_check


def _rel_elem_floor(got, want, floor_frac=1e-3):
    """Element-wise relative error with a floor: max |got - want| / max(|want|, floor_frac * max|want|).  A voxel whose
    gradient is 0.1 % of the batch maximum still has to be right to ELEM_TOL of its own size."""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    floor = floor_frac * float(np.max(np.abs(want)))
    return float(np.max(np.abs(got - want) / np.maximum(np.abs(want), floor)))


def _elem_check(got, want64, ref32):
    """Element-wise gradient bar.  Every element must agree with the float64 oracle to ELEM_TOL of max(|element|,
    0.1 % of the batch maximum) -- or, where float32 arithmetic itself cannot deliver that, to within twice the distance
    of the REFERENCE's own float32 evaluation (``ref32``: the reference-source fixture or the float32-emulated oracle)
    from the same float64 values.  Measured: the reference's float32 gradients sit 1.4e-3 from float64 in this metric
    (absolute error 2e-5 on gradients of size 13), so a flat 1e-4 would fail the reference itself."""
    e = _rel_elem_floor(got, want64)
    bar = max(ELEM_TOL, 2.0 * _rel_elem_floor(ref32, want64))
    assert e < bar, 'element-wise error %.3g exceeds %.3g (reference float32 vs float64: %.3g)' % (
        e, bar, _rel_elem_floor(ref32, want64))
    # entries of at least 10 % of the batch maximum meet the flat tolerance unconditionally
    assert _rel_elem_floor(got, want64, 1e-1) < ELEM_TOL


@pytest.fixture(scope='module')
def qb():
    import qbold_vi_b200 as qb
    qb._lib.lib()
    return qb


@pytest.fixture(scope='module')
def dev():
    return torch.device('cuda', 0)


def _t(a, dev):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32), device=dev)


def _rand_voxels(n, seed=0):
    rng = np.random.default_rng(seed)
    return np.stack([rng.uniform(0.04, 0.84, n), rng.uniform(0.001, 0.201, n)], -1).astype(np.float32)


# ---------------------------------------------------------------------------------- forward
@pytest.mark.parametrize('full', [True, False])
@pytest.mark.parametrize('blood', [True, False])
def test_forward_matches_oracle(qb, dev, cfg_noise_off, physics, full, blood):
    x = _rand_voxels(4096, 1)
    layer = qb.SignalGenerationLayer(cfg_noise_off, full, blood)
    s = layer(_t(x, dev)).cpu().numpy()
    assert s.shape == (4096, 11) and s.dtype == np.float32
    s32 = o.forward(physics, x, full, blood, np.float32)
    s64 = o.forward(physics, x, full, blood, np.float64)
    assert rel_elem(s, s32) < SIG_TOL, 'vs float32-emulated TF semantics'
    assert rel_elem(s, s64) < SIG_TOL, 'vs float64 restatement (node 0 dead)'


def test_forward_matches_reference_source_fixture(qb, dev, cfg_noise_off):
    g = golden('ref_shim_forward.npz')
    for full in (1, 0):
        for blood in (1, 0):
            layer = qb.SignalGenerationLayer(cfg_noise_off, bool(full), bool(blood), taus=g['taus'])
            s = layer(_t(g['oef_dbv'], dev).reshape(-1, 1, 1, 1, 2))
            assert tuple(s.shape) == (g['oef_dbv'].shape[0], 1, 1, 1, 11)          # leading shape kept (signals.py:137-140)
            assert rel_elem(s.cpu().numpy().reshape(-1, 11), g['signal_f%d_b%d' % (full, blood)]) < SIG_TOL


def test_appendix_b_known_answers(qb, dev, cfg_noise_off):
    k = golden('kat_appendix_b.npz')
    layer = qb.SignalGenerationLayer(cfg_noise_off, True, True)
    s = layer(_t(k['oef_dbv'], dev)).cpu().numpy()
    assert np.max(np.abs(s[0] - k['signal_fp64_0'])) < 2e-6
    assert np.max(np.abs(s[1] - k['signal_fp32_1'])) < 2e-6
    _, g = layer.forward_backward(_t(k['oef_dbv'][:1], dev))                        # g_signal NULL = ones (signals.py:307-314 demo)
    assert rel_elem(g.cpu().numpy()[0], k['grad_sum_0']) < GRAD_TOL


@pytest.mark.parametrize('full', [True, False])
@pytest.mark.parametrize('blood', [True, False])
def test_forward_backward_matches_tf_consistent_gradient(qb, dev, cfg_noise_off, physics, full, blood):
    x = _rand_voxels(2048, 2)
    gs = np.random.default_rng(5).standard_normal((2048, 11)).astype(np.float32)
    layer = qb.SignalGenerationLayer(cfg_noise_off, full, blood)
    s, g = layer.forward_backward(_t(x, dev), _t(gs, dev))
    s64, g64 = o.forward_backward(physics, x, gs, full, blood, np.float64)
    assert rel_elem(s.cpu().numpy(), s64) < SIG_TOL
    g = g.cpu().numpy()
    assert rel_max(g[:, 0], g64[:, 0]) < GRAD_TOL and rel_max(g[:, 1], g64[:, 1]) < GRAD_TOL
    # element-wise, away from sign changes of the random-weighted sum
    big = np.abs(g64) > 1e-2 * np.abs(g64).max(0, keepdims=True)
    assert np.max(np.abs(g - g64)[big] / np.abs(g64)[big]) < 5 * GRAD_TOL
    # the golden gradient of the reference source (tape.gradient over the TF shim)
    gold = golden('ref_shim_forward.npz')
    layer = qb.SignalGenerationLayer(cfg_noise_off, full, blood, taus=gold['taus'])
    key = 'f%d_b%d' % (int(full), int(blood))
    _, g1 = layer.forward_backward(_t(gold['oef_dbv'], dev))
    _, g2 = layer.forward_backward(_t(gold['oef_dbv'], dev), _t(gold['g_rand'], dev))
    assert rel_max(g1.cpu().numpy(), gold['grad_ones_' + key]) < GRAD_TOL
    assert rel_max(g2.cpu().numpy(), gold['grad_rand_' + key]) < GRAD_TOL


def test_autograd_through_the_layer(qb, dev, cfg_noise_off):
    x = _t(_rand_voxels(257, 3), dev).reshape(257, 1, 2).requires_grad_(True)
    w = torch.randn(257, 1, 11, device=dev)
    layer = qb.SignalGenerationLayer(cfg_noise_off, True, True)
    (layer(x) * w).sum().backward()
    _, g = layer.forward_backward(x.detach(), w)
    assert torch.equal(x.grad.reshape(-1, 2), g)


def test_edge_cases(qb, dev, cfg_noise_off, physics):
    layer = qb.SignalGenerationLayer(cfg_noise_off, True, True)
    assert tuple(layer(torch.zeros((0, 2), device=dev)).shape) == (0, 11)          # empty input
    for n in (1, 7, 33, 255, 1025):                                                 # ragged sizes
        x = _rand_voxels(n, n)
        assert rel_elem(layer(_t(x, dev)).cpu().numpy(), o.forward(physics, x, dtype=np.float64)) < SIG_TOL
    with pytest.raises(AssertionError):
        layer(torch.zeros((4, 3), device=dev))
    # corners of the admissible box and beyond it (OEF up to 1: a = 29)
    x = np.array([[0.04, 0.001], [0.84, 0.201], [0.04, 0.201], [0.84, 0.001], [1.0, 0.3], [0.01, 0.0005]], np.float32)
    assert rel_elem(layer(_t(x, dev)).cpu().numpy(), o.forward(physics, x, dtype=np.float32)) < SIG_TOL
    # variable haematocrit input (signals.py:64-70)
    xv = np.concatenate([_rand_voxels(300, 9), np.random.default_rng(1).uniform(0.25, 0.45, (300, 1))], -1).astype(np.float32)
    lv = qb.SignalGenerationLayer(cfg_noise_off, True, True, variable_hct=True)
    assert rel_elem(lv(_t(xv, dev)).cpu().numpy(), o.forward(physics, xv, variable_hct=True, dtype=np.float64)) < SIG_TOL


def test_24_tau_grid(qb, dev, cfg_noise_off):
    cfg = dict(cfg_noise_off, tau_start='-0.028', tau_end='0.065', tau_step='0.004')
    ph = o.parse_params(cfg)
    layer = qb.SignalGenerationLayer(cfg, True, True)
    assert layer.params.n_cols == 16                                                # two column groups
    x = _rand_voxels(512, 4)
    gs = np.random.default_rng(6).standard_normal((512, 24)).astype(np.float32)
    s, g = layer.forward_backward(_t(x, dev), _t(gs, dev))
    s64, g64 = o.forward_backward(ph, x, gs, dtype=np.float64)
    assert rel_elem(s.cpu().numpy(), s64) < SIG_TOL and rel_max(g.cpu().numpy(), g64) < GRAD_TOL


# ---------------------------------------------------------------------------------- sampling / likelihood / KL
def _trainer(qb, cfg, e=None, **kw):
    args = dict(student_t_df=200, multi_image_normalisation=False, use_mvg=True, use_population_prior=False,
                predict_log_data=False, seed=1234)
    args.update(kw)
    return qb.EncoderTrainer(cfg, **args)


def test_reparam_layer_matches_reference_source(qb, dev, cfg_noise_off):
    e = golden('ref_shim_elbo_optimal.npz')
    tr = _trainer(qb, cfg_noise_off)
    q = _t(e['q'], dev).reshape(2, 4, 4, 2, 5).requires_grad_(True)
    smp = qb.ReparamTrickLayer(tr)((q, None), eps=_t(e['eps'], dev))
    assert tuple(smp.shape) == (2, 4, 4, 2, 2)
    assert rel_elem(smp.detach().cpu().numpy().reshape(-1, 2), e['sampled']) < SIG_TOL
    # its gradient against torch autograd of the same formulas
    w = torch.randn_like(smp)
    (smp * w).sum().backward()
    q2 = _t(e['q'], dev).requires_grad_(True)
    ep = _t(e['eps'], dev)
    z_o = q2[:, 0] + ep[:, 0] * torch.exp(tr.transform_std(q2[:, 1]))
    z_d = q2[:, 2] + ep[:, 0] * tr.transform_offdiag(q2[:, 4]) + ep[:, 1] * torch.exp(tr.transform_std(q2[:, 3]))
    ref = tr.forward_transform(torch.stack([z_o, z_d], -1))
    (ref * w.reshape(-1, 2)).sum().backward()
    assert rel_max(q.grad.cpu().numpy().reshape(-1, 5), q2.grad.cpu().numpy()) < 1e-5


@pytest.mark.parametrize('tag', ['optimal', 'multinorm', 'studentt'])
def test_fused_elbo_matches_reference_source(qb, dev, cfg_noise_off, tag):
    e = golden('ref_shim_elbo_%s.npz' % tag)
    df = float(e['student_t_df'])
    tr = _trainer(qb, cfg_noise_off, student_t_df=df, multi_image_normalisation=bool(e['multi_image_normalisation']))
    layer = qb.SignalGenerationLayer(cfg_noise_off, True, True)
    q = _t(e['q'], dev).requires_grad_(True)
    sg = _t(e['sigma'], dev).requires_grad_(True)
    loss, info = tr.fused_elbo(layer, q, sg, _t(e['data'], dev), _t(e['mask'], dev), _t(e['prior'], dev),
                               kl_samples=70, eps=_t(e['eps'], dev), eps_kl=_t(e['eps_kl'], dev), return_maps=True)
    loss.backward()
    assert rel_elem(info['nll'].item(), e['nll']) < GRAD_TOL
    assert rel_elem(info['kl'].item(), e['kl']) < GRAD_TOL
    assert rel_elem(loss.item(), e['nll'] + e['kl']) < GRAD_TOL
    assert rel_max(info['nll_map'].cpu().numpy(), e['nll_map']) < GRAD_TOL
    assert rel_max(q.grad.cpu().numpy(), e['grad_q_nll'] + e['grad_q_kl']) < GRAD_TOL
    assert rel_max(sg.grad.cpu().numpy(), e['grad_sigma']) < GRAD_TOL
    if tag != 'studentt':                                     # the oracle's analytic backward covers the Gaussian likelihood
        r64 = o.elbo_and_grads(o.parse_params(cfg_noise_off), e['q'], e['sigma'], e['data'], e['mask'], e['prior'], e['eps'],
                               e['eps_kl'], np.float64, multi_image_normalisation=bool(e['multi_image_normalisation']))
        _elem_check(q.grad.cpu().numpy(), r64['grad_q'], e['grad_q_nll'] + e['grad_q_kl'])
        _elem_check(sg.grad.cpu().numpy(), r64['grad_sigma'], e['grad_sigma'])
    assert float(info['mask_sum']) == e['mask'].sum() and float(info['non_finite']) == 0
    # masked voxels: exactly zero loss and gradient (model.py:564,661)
    dead = e['mask'] == 0
    assert np.all(q.grad.cpu().numpy()[dead] == 0) and np.all(sg.grad.cpu().numpy()[dead] == 0)


def test_fused_elbo_matches_oracle_on_random_batch(qb, dev, cfg_noise_off, physics):
    n, S = 1536, 70
    r = np.random.default_rng(21)
    q = np.stack([r.normal(-0.3, 0.7, n), r.normal(0, 0.6, n), r.normal(-1.2, 0.7, n), r.normal(0, 0.6, n),
                  r.normal(0, 0.8, n)], -1).astype(np.float32)
    prior = (q + r.normal(0, 0.3, (n, 5))).astype(np.float32)
    sigma = np.exp(r.normal(np.log(0.05), 0.2, (n, 11))).astype(np.float32)
    mask = (r.uniform(size=n) > 0.3).astype(np.float32)
    truth = np.stack([r.uniform(0.1, 0.7, n), r.uniform(0.005, 0.15, n)], -1)
    data = (o.forward(physics, truth, dtype=np.float64) * 100 * (1 + 0.02 * r.standard_normal((n, 11)))).astype(np.float32)
    data *= mask[:, None]
    eps = r.standard_normal((n, 2)).astype(np.float32)
    eps_kl = r.standard_normal((n, S, 2)).astype(np.float32)
    ref = o.elbo_and_grads(physics, q, sigma, data, mask, prior, eps, eps_kl, np.float64)
    tr = _trainer(qb, cfg_noise_off)
    layer = qb.SignalGenerationLayer(cfg_noise_off, True, True)
    qt, st = _t(q, dev).requires_grad_(True), _t(sigma, dev).requires_grad_(True)
    loss, info = tr.fused_elbo(layer, qt, st, _t(data, dev), _t(mask, dev), _t(prior, dev), kl_samples=S,
                               eps=_t(eps, dev), eps_kl=_t(eps_kl, dev), return_maps=True)
    loss.backward()
    assert rel_elem(info['nll'].item(), ref['nll']) < GRAD_TOL and rel_elem(info['kl'].item(), ref['kl']) < GRAD_TOL
    assert rel_max(info['kl_map'].cpu().numpy(), ref['kl_map']) < GRAD_TOL
    assert rel_max(qt.grad.cpu().numpy(), ref['grad_q']) < GRAD_TOL
    assert rel_max(st.grad.cpu().numpy(), ref['grad_sigma']) < GRAD_TOL
    ref32 = o.elbo_and_grads(physics, q, sigma, data, mask, prior, eps, eps_kl, np.float32)
    _elem_check(qt.grad.cpu().numpy(), ref['grad_q'], ref32['grad_q'])
    _elem_check(st.grad.cpu().numpy(), ref['grad_sigma'], ref32['grad_sigma'])
    # sharding invariance: the same batch in two halves with the global mask sum gives the same gradients
    h = n // 2
    parts = []
    for sl in (slice(0, h), slice(h, n)):
        qh = _t(q[sl], dev).requires_grad_(True)
        l, _ = tr.fused_elbo(layer, qh, _t(sigma[sl], dev), _t(data[sl], dev), _t(mask[sl], dev), _t(prior[sl], dev),
                             kl_samples=S, eps=_t(eps[sl], dev), eps_kl=_t(eps_kl[sl], dev), mask_sum=mask.sum())
        l.backward()
        parts.append((l.item(), qh.grad))
    assert abs(parts[0][0] + parts[1][0] - loss.item()) < 1e-5 * abs(loss.item())
    assert torch.equal(torch.cat([parts[0][1], parts[1][1]]), qt.grad)


def _elbo_batch(physics, n, seed, spread=0.7):
    r = np.random.default_rng(seed)
    q = np.stack([r.normal(-0.3, spread, n), r.normal(0, 0.6, n), r.normal(-1.2, spread, n), r.normal(0, 0.6, n),
                  r.normal(0, 0.8, n)], -1).astype(np.float32)
    prior = (q + r.normal(0, 0.3, (n, 5))).astype(np.float32)
    sigma = np.exp(r.normal(np.log(0.05), 0.2, (n, 11))).astype(np.float32)
    truth = np.stack([r.uniform(0.1, 0.7, n), r.uniform(0.005, 0.15, n)], -1)
    data = (o.forward(physics, truth, dtype=np.float64) * 100 * (1 + 0.02 * r.standard_normal((n, 11)))).astype(np.float32)
    return q, prior, sigma, data


def test_production_rng_path_matches_oracle(qb, dev, cfg_noise_off, physics):
    """The calls as users make them -- fused_elbo(seed=...) and kl_loss() with NO explicit draws, i.e. in-kernel
    Philox4x32-10 + Box-Muller -- against the oracle fed oracle/philox.py draws for the same (seed, global index)."""
    n, S, seed, off = 2048, 70, 99, 1_234_567
    q, prior, sigma, data = _elbo_batch(physics, n, 31)
    mask = (np.random.default_rng(32).uniform(size=n) > 0.3).astype(np.float32)
    data *= mask[:, None]
    idx = np.arange(n, dtype=np.uint64) + np.uint64(off)
    eps, eps_kl = philox.reparam_eps(seed, idx), philox.kl_eps(seed, idx, S)
    ref = o.elbo_and_grads(physics, q, sigma, data, mask, prior, eps, eps_kl, np.float64)
    tr = _trainer(qb, cfg_noise_off)
    layer = qb.SignalGenerationLayer(cfg_noise_off, True, True)
    qt, st = _t(q, dev).requires_grad_(True), _t(sigma, dev).requires_grad_(True)
    loss, info = tr.fused_elbo(layer, qt, st, _t(data, dev), _t(mask, dev), _t(prior, dev), kl_samples=S, seed=seed,
                               offset=off, return_maps=True)
    loss.backward()
    assert rel_elem(info['nll'].item(), ref['nll']) < GRAD_TOL and rel_elem(info['kl'].item(), ref['kl']) < GRAD_TOL
    assert rel_elem(loss.item(), ref['nll'] + ref['kl']) < GRAD_TOL
    assert rel_max(info['nll_map'].cpu().numpy(), ref['nll_map']) < GRAD_TOL
    assert rel_max(info['kl_map'].cpu().numpy(), ref['kl_map']) < GRAD_TOL
    ref32 = o.elbo_and_grads(physics, q, sigma, data, mask, prior, eps, eps_kl, np.float32)
    for got, key in ((qt.grad, 'grad_q'), (st.grad, 'grad_sigma')):
        assert rel_max(got.cpu().numpy(), ref[key]) < GRAD_TOL
        _elem_check(got.cpu().numpy(), ref[key], ref32[key])
    # kl_loss() with its defaults: the trainer's own call counter seeds the draws
    tr2 = _trainer(qb, cfg_noise_off, seed=77)
    seed2 = (77 + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF                           # first call of this trainer
    true = torch.cat([_t(prior, dev), _t(mask, dev)[:, None]], -1)
    pred = _t(q, dev).requires_grad_(True)
    kl = tr2.kl_loss(true, pred)
    kl.backward()
    eps_kl2 = philox.kl_eps(seed2, np.arange(n, dtype=np.uint64), 70)
    want = float(np.sum(o.mc_kl(prior, q, eps_kl2, np.float64) * mask) / mask.sum())
    assert rel_elem(kl.item(), want) < GRAD_TOL
    ref2 = o.elbo_and_grads(physics, q, sigma, data, mask, prior, np.zeros((n, 2), np.float32), eps_kl2, np.float64)
    assert rel_elem(kl.item(), ref2['kl']) < GRAD_TOL
    g_kl = pred.grad.cpu().numpy()
    assert rel_max(g_kl, ref2['grad_q_kl']) < GRAD_TOL
    ref2_32 = o.elbo_and_grads(physics, q, sigma, data, mask, prior, np.zeros((n, 2), np.float32), eps_kl2, np.float32)
    _elem_check(g_kl, ref2['grad_q_kl'], ref2_32['grad_q_kl'])


def test_kl_loss_mc_and_closed_form(qb, dev, cfg_noise_off):
    e = golden('ref_shim_elbo_optimal.npz')
    tr = _trainer(qb, cfg_noise_off)
    true = torch.cat([_t(e['prior'], dev), _t(e['mask'], dev)[:, None]], -1).reshape(2, 4, 4, 2, 6)
    pred = _t(e['q'], dev).reshape(2, 4, 4, 2, 5).requires_grad_(True)
    kl = tr.kl_loss(true, pred, eps=_t(e['eps_kl'], dev))
    kl.backward()
    assert rel_elem(kl.item(), e['kl']) < GRAD_TOL
    assert rel_max(pred.grad.cpu().numpy().reshape(-1, 5), e['grad_q_kl']) < GRAD_TOL
    kl_map = tr.kl_loss(true, pred.detach(), return_mean=False, eps=_t(e['eps_kl2'], dev))
    assert tuple(kl_map.shape) == (2, 4, 4, 2, 1)
    assert rel_max(kl_map.cpu().numpy().reshape(-1), e['kl_map2']) < GRAD_TOL
    # closed form: value vs the textbook formula, gradient vs float64 central differences of it
    p2 = _t(e['q'], dev).requires_grad_(True)
    klc = tr.kl_loss(true.reshape(-1, 6), p2, return_mean=False, no_samples=0)
    klc.sum().backward()
    ref = np.where(e['mask'] > 0, o.closed_form_kl(e['prior'], e['q']), 0)
    assert rel_max(klc.detach().cpu().numpy().reshape(-1), ref) < 1e-5
    h = 1e-6
    for j in range(5):
        d = np.zeros(5)
        d[j] = h
        fd = (o.closed_form_kl(e['prior'], e['q'].astype(np.float64) + d) -
              o.closed_form_kl(e['prior'], e['q'].astype(np.float64) - d)) / (2 * h)
        fd = np.where(e['mask'] > 0, fd, 0)
        assert rel_max(p2.grad.cpu().numpy()[:, j], fd) < GRAD_TOL


def test_posterior_stats(qb, dev, cfg_noise_off, physics):
    m = golden('ref_shim_means.npz')
    tr = _trainer(qb, cfg_noise_off)
    q = _t(m['q'], dev).reshape(2, 4, 4, 2, 5)
    means, var = tr.calculate_means(q, torch.ones_like(q[..., :1]), include_r2p=True, return_stds=True, no_samples=16,
                                    eps=_t(m['eps'], dev))
    assert tuple(means.shape) == (2, 4, 4, 2, 3)
    assert rel_elem(means.cpu().numpy().reshape(-1, 3), m['means']) < SIG_TOL
    assert rel_max(var.cpu().numpy().reshape(-1, 3), m['stds']) < GRAD_TOL         # 'stds' are variances (model.py:331)
    # in-kernel Philox draws == oracle/philox.py draws (64 samples, BASELINE config 4)
    tr2 = _trainer(qb, cfg_noise_off, seed=77)
    seed = (77 + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF                           # first call of this trainer
    mp, vp = tr2.calculate_means(q, None, include_r2p=True, return_stds=True, no_samples=64)
    eps = philox.kl_eps(seed, np.arange(64), 64)
    mo, vo = o.posterior_stats(physics, m['q'], eps, np.float64)
    assert rel_elem(mp.cpu().numpy().reshape(-1, 3), mo) < GRAD_TOL
    assert rel_max(vp.cpu().numpy().reshape(-1, 3), vo) < 10 * GRAD_TOL


# ---------------------------------------------------------------------------------- synthetic generation
@pytest.mark.parametrize('tag', ['u10', 'u0'])
def test_generation_matches_reference_source(qb, dev, tag):
    d = golden('ref_shim_dataset_%s.npz' % tag)
    cfg = o.default_config()                                                         # noise ON, as in the INI
    ty = d['train_y']
    # un-shuffle the labels to recover the marginals the reference drew
    inv = np.empty(529, dtype=np.int64)
    inv[np.arange(529)] = d['perm']
    grid = np.empty((529, 2), np.float32)
    grid[d['perm']] = ty[:, :2]
    oefs, dbvs = grid.reshape(23, 23, 2)[:, 0, 0].copy(), grid.reshape(23, 23, 2)[0, :, 1].copy()
    layer = qb.SignalGenerationLayer(cfg, True, True)
    snr_u01 = _t(d['snr_u01'].reshape(-1), dev)
    x, y = qb.generate_from_marginals(layer, _t(oefs, dev), _t(dbvs, dev),
                                      torch.as_tensor(d['perm'], device=dev).contiguous(), n_chunks=10,
                                      snr_u01=snr_u01, noise_eps=_t(d['noise_eps'], dev))
    assert tuple(x.shape) == (520, 11) and tuple(y.shape) == (529, 3)               # S^2 % 10 rows dropped from x only
    assert rel_elem(y.cpu().numpy(), ty) < 1e-6
    assert rel_elem(x.cpu().numpy(), d['train_x']) < 2 * SIG_TOL


def test_generation_feistel_and_philox_noise(qb, dev, physics):
    cfg = o.default_config()
    ph = o.parse_params(cfg)
    layer = qb.SignalGenerationLayer(cfg, True, True, seed=4242)
    r = np.random.default_rng(8)
    oefs = r.uniform(0.05, 0.8, 40).astype(np.float32)
    dbvs = r.uniform(0.003, 0.195, 30).astype(np.float32)
    x, y = qb.generate_from_marginals(layer, _t(oefs, dev), _t(dbvs, dev), None, n_chunks=10, seed=4242)
    perm = philox.feistel_permute(np.arange(1200), 1200, 4242)
    assert sorted(perm.tolist()) == list(range(1200))
    snr = np.concatenate([philox.snr_u01(4242 ^ 0x5DEECE66D, np.arange(i * 120, (i + 1) * 120)) for i in range(10)])
    eps = np.concatenate([philox.noise_eps(4242 ^ 0x5DEECE66D, np.arange(i * 120, (i + 1) * 120), 11) for i in range(10)])
    xo, yo = o.synthetic_dataset_from_draws(ph, oefs, dbvs, perm, snr * np.float32(70) + np.float32(50), eps)
    assert rel_elem(y.cpu().numpy(), yo) < 1e-6
    assert rel_elem(x.cpu().numpy(), xo) < 5 * SIG_TOL
    # public entry point: shapes, label ranges, noise level
    cfg['sample_size'] = '100'
    tx, ty = qb.create_synthetic_dataset(cfg, True, True, 0.0, uniform_prop=0.1, device=dev, seed=3)
    assert tuple(tx.shape) == (10000, 11) and tuple(ty.shape) == (10000, 3)
    ty = ty.cpu().numpy()
    assert ty[:, 0].min() >= 0.05 - 1e-6 and ty[:, 0].max() <= 0.8 + 1e-6
    assert ty[:, 1].min() >= 0.003 - 1e-6 and ty[:, 1].max() <= 0.195 + 1e-6
    clean = o.forward(ph, ty[:, :2], dtype=np.float32)
    resid = tx.cpu().numpy() - clean
    snr_eff = clean.mean(0) / resid.std(0)
    assert np.all(snr_eff > 40) and np.all(snr_eff < 130)                           # U(50,120) * norm_snr
    cfg5 = dict(cfg, tau_start='0.0', tau_end='0.05', tau_step='0.01')
    with pytest.raises(UnboundLocalError):                                          # signals.py:117-121
        qb.SignalGenerationLayer(cfg5, True, True)(torch.rand(8, 2, device=dev) * 0.1 + 0.1)


# ---------------------------------------------------------------------------------- full-size properties
def test_full_size_properties_16M(qb, dev, cfg_noise_off):
    """BASELINE config 2 size: 16 777 216 voxels.  Size-independent properties only."""
    n = 1 << 24
    g = torch.Generator(device=dev).manual_seed(1234)
    x = torch.rand((n, 2), device=dev, generator=g)
    x[:, 0] = x[:, 0] * 0.8 + 0.04
    x[:, 1] = x[:, 1] * 0.2 + 0.001
    tissue = qb.SignalGenerationLayer(cfg_noise_off, True, False)                    # S = (1 - dbv) * S_t
    s = tissue(x)
    assert torch.isfinite(s).all()
    assert torch.equal(s[:, 0], s[:, 4]) and torch.equal(s[:, 1], s[:, 3])           # S(tau) = S(-tau)
    e = float(np.float32(tissue.params.e_tissue))
    assert torch.allclose(s[:, 2], (1 - x[:, 1]) * e, rtol=3e-7, atol=0)             # tau = 0: exp(-TE*R2t)
    assert (s[:, 5:] <= s[:, 4:-1] * (1 + 1e-6)).all()                               # decays with tau
    layer = qb.SignalGenerationLayer(cfg_noise_off, True, True)
    sl = slice(5_000_000, 5_000_000 + 4096)
    w1 = torch.randn((n, 11), device=dev, generator=g)
    s_full, g1 = layer.forward_backward(x, w1)
    s_part, g_part = layer.forward_backward(x[sl].contiguous(), w1[sl].contiguous())
    assert torch.equal(s_full[sl], s_part) and torch.equal(g1[sl], g_part)          # shard invariance, bit-exact
    # oracle parity AT full size: a random 100 000-voxel subsample of the 16 M outputs vs the float64 oracle
    pick = torch.randperm(n, device=dev, generator=g)[:100_000]
    ph = o.parse_params(cfg_noise_off)
    s64, g64 = o.forward_backward(ph, x[pick].cpu().numpy(), w1[pick].cpu().numpy(), dtype=np.float64)
    assert rel_elem(s_full[pick].cpu().numpy(), s64) < SIG_TOL
    assert rel_max(g1[pick].cpu().numpy(), g64) < GRAD_TOL
    _, g32 = o.forward_backward(ph, x[pick[:20_000]].cpu().numpy(), w1[pick[:20_000]].cpu().numpy(), dtype=np.float32)
    _elem_check(g1[pick[:20_000]].cpu().numpy(), g64[:20_000], g32)
    del s, s_full
    _, g_ones = layer.forward_backward(x, None, want_signal=False)
    _, g2 = layer.forward_backward(x, w1 + 1.0, want_signal=False)
    err = (g2 - (g1 + g_ones)).abs().max(0).values / g2.abs().max(0).values         # the VJP is linear in g
    assert float(err.max()) < 1e-5


# ---------------------------------------------------------------------------------- training glue
def test_config3_volume_elbo_matches_oracle(qb, dev, cfg_noise_off, physics):
    """BASELINE config 3 shape: one 64^3 volume, sphere mask r = 28 (92 k live voxels of 262 144), 70-sample KL with
    the in-kernel Philox draws, against the float64 oracle fed the same draws."""
    side, S, seed = 64, 70, 2026
    n = side ** 3
    q, prior, sigma, data = _elbo_batch(physics, n, 64, spread=0.5)
    zz, yy, xx = np.meshgrid(*(np.arange(side) - (side - 1) / 2.0,) * 3, indexing='ij')
    mask = ((xx ** 2 + yy ** 2 + zz ** 2) <= 28.0 ** 2).astype(np.float32).reshape(n)
    data *= mask[:, None]
    idx = np.arange(n, dtype=np.uint64)
    ref = o.elbo_and_grads(physics, q, sigma, data, mask, prior, philox.reparam_eps(seed, idx),
                           philox.kl_eps(seed, idx, S), np.float64)
    tr = _trainer(qb, cfg_noise_off)
    layer = qb.SignalGenerationLayer(cfg_noise_off, True, True)
    shp = (1, side, side, side)
    qt = _t(q, dev).reshape(shp + (5,)).requires_grad_(True)
    st = _t(sigma, dev).reshape(shp + (11,)).requires_grad_(True)
    loss, info = tr.fused_elbo(layer, qt, st, _t(data, dev).reshape(shp + (11,)), _t(mask, dev).reshape(shp + (1,)),
                               _t(prior, dev).reshape(shp + (5,)), kl_samples=S, seed=seed, return_maps=True)
    loss.backward()
    assert float(info['mask_sum']) == mask.sum()
    assert rel_elem(info['nll'].item(), ref['nll']) < GRAD_TOL and rel_elem(info['kl'].item(), ref['kl']) < GRAD_TOL
    assert rel_max(info['nll_map'].cpu().numpy().reshape(-1), ref['nll_map']) < GRAD_TOL
    assert rel_max(info['kl_map'].cpu().numpy().reshape(-1), ref['kl_map']) < GRAD_TOL
    gq, gs = qt.grad.cpu().numpy().reshape(n, 5), st.grad.cpu().numpy().reshape(n, 11)
    assert rel_max(gq, ref['grad_q']) < GRAD_TOL and rel_max(gs, ref['grad_sigma']) < GRAD_TOL
    sub = np.flatnonzero(mask > 0)[::5]                                              # float32 reference on 18 k live voxels
    ref32 = o.elbo_and_grads(physics, q[sub], sigma[sub], data[sub], mask[sub], prior[sub],
                             philox.reparam_eps(seed, idx[sub]), philox.kl_eps(seed, idx[sub], S), np.float32)
    scale = float(mask[sub].sum() / mask.sum())                                      # the oracle divides by ITS sum(mask)
    _elem_check(gq[sub], ref['grad_q'][sub], ref32['grad_q'] * scale)
    _elem_check(gs[sub], ref['grad_sigma'][sub], ref32['grad_sigma'] * scale)
    assert np.all(gq[mask == 0] == 0) and np.all(gs[mask == 0] == 0)


def test_fused_step_equals_unfused_layer_composition(qb, dev, cfg_noise_off):
    """One fused launch == the reference's graph built from the separate layers (build_fine_tuner,
    model.py:239-286 + loss closures train.py:315-320), gradients w.r.t. the encoder parameters included."""
    from qbold_vi_b200.encoder import Encoder
    torch.manual_seed(0)
    enc = Encoder(no_units=16, no_intermediate_layers=1).to(dev)
    tr = _trainer(qb, cfg_noise_off)
    layer = qb.SignalGenerationLayer(cfg_noise_off, True, True)
    g = torch.Generator(device=dev).manual_seed(5)
    shape = (2, 6, 5, 3)
    truth = torch.stack([torch.rand(shape, device=dev, generator=g) * 0.5 + 0.15,
                         torch.rand(shape, device=dev, generator=g) * 0.1 + 0.01], -1)
    mask = (torch.rand(shape + (1,), device=dev, generator=g) > 0.25).float()
    data = layer(truth) * 100.0 * (1 + 0.02 * torch.randn(shape + (11,), device=dev, generator=g)) * mask
    prior = torch.randn(shape + (5,), device=dev, generator=g) * 0.4
    n = mask.numel()
    eps = torch.randn((n, 2), device=dev, generator=g)
    eps_kl = torch.randn((n, 70, 2), device=dev, generator=g)

    _, q, sigma = enc(data)
    loss_f, info = tr.fused_elbo(layer, q, sigma, data, mask, prior, kl_samples=70, eps=eps, eps_kl=eps_kl)
    g_f = torch.autograd.grad(loss_f, list(enc.parameters()))

    _, q, sigma = enc(data)
    sampled = qb.ReparamTrickLayer(tr)((q, mask), eps=eps)                           # model.py:248
    pred = layer(sampled)                                                            # model.py:273
    nll = tr.fine_tune_loss_fn(torch.cat([data, mask], -1), torch.cat([pred, sigma], -1))
    kl = tr.kl_loss(torch.cat([prior, mask], -1), q, eps=eps_kl)
    g_u = torch.autograd.grad(nll + kl, list(enc.parameters()))
    assert abs(loss_f.item() - (nll + kl).item()) < GRAD_TOL * abs(loss_f.item())
    assert abs(info['nll'].item() - nll.item()) < GRAD_TOL * abs(nll.item())
    for a, b in zip(g_f, g_u):
        assert rel_max(a.cpu().numpy(), b.cpu().numpy()) < 5 * GRAD_TOL


def test_single_gpu_training_steps_reduce_the_loss(qb, dev, cfg_noise_off):
    from qbold_vi_b200.encoder import Encoder
    from qbold_vi_b200.distributed import DataParallelTrainer
    torch.manual_seed(1)
    enc = Encoder().to(dev)
    tr = _trainer(qb, cfg_noise_off)
    layer = qb.SignalGenerationLayer(cfg_noise_off, True, True)
    dp = DataParallelTrainer(enc, tr, layer, ft_lr=2e-3)
    g = torch.Generator(device=dev).manual_seed(9)
    shape = (2, 16, 16, 4)
    truth = torch.stack([torch.rand(shape, device=dev, generator=g) * 0.4 + 0.2,
                         torch.rand(shape, device=dev, generator=g) * 0.06 + 0.01], -1)
    mask = torch.ones(shape + (1,), device=dev)
    data = layer(truth) * 100.0
    with torch.no_grad():
        prior = enc(data)[0].clone()
    losses = [dp.step(data, mask, prior) for _ in range(12)]
    assert all(np.isfinite(s['loss']) for s in losses)
    assert np.mean([s['nll'] for s in losses[-3:]]) < np.mean([s['nll'] for s in losses[:3]])
    # the step itself never synchronises with the host: torch raises on any synchronising call in this mode (the mask
    # count, the losses and the statistics stay on the device until somebody reads them)
    torch.cuda.synchronize()
    torch.cuda.set_sync_debug_mode('error')
    try:
        lazy = [dp.step(data, mask, prior) for _ in range(3)]
    finally:
        torch.cuda.set_sync_debug_mode('default')
    assert np.isfinite(lazy[-1]['loss']) and lazy[-1]['mask_sum'] == float(mask.sum())


def test_captured_training_step_matches_the_eager_step(qb, dev, cfg_noise_off, tmp_path):
    """cuda_graph=True: the step is captured after three eager warm-up steps and replayed; the Philox key, the
    schedule position and Adam's counter advance on the device.  Same seeds, same data -> the statistics of every
    step (warm-up, first replay, later replays, after a checkpoint round trip) equal the eager trainer's."""
    _captured_step_against_eager(qb, dev, cfg_noise_off, tmp_path, True)


def _captured_step_against_eager(qb, dev, cfg_noise_off, tmp_path, graph_mode):
    import copy
    from qbold_vi_b200.encoder import Encoder
    from qbold_vi_b200.distributed import DataParallelTrainer
    torch.manual_seed(1)
    enc_e = Encoder().to(dev)
    enc_g = copy.deepcopy(enc_e)
    layer = qb.SignalGenerationLayer(cfg_noise_off, True, True)
    dp_e = DataParallelTrainer(enc_e, _trainer(qb, cfg_noise_off, seed=11), layer, ft_lr=2e-3)
    dp_g = DataParallelTrainer(enc_g, _trainer(qb, cfg_noise_off, seed=11), layer, ft_lr=2e-3, cuda_graph=graph_mode)
    g = torch.Generator(device=dev).manual_seed(9)
    shape = (2, 16, 16, 4)
    truth = torch.stack([torch.rand(shape, device=dev, generator=g) * 0.4 + 0.2,
                         torch.rand(shape, device=dev, generator=g) * 0.06 + 0.01], -1)
    mask = (torch.rand(shape + (1,), device=dev, generator=g) > 0.2).float()
    data = layer(truth) * 100.0 * mask
    with torch.no_grad():
        prior = enc_e(data)[0].clone()

    def close(a, b, what):
        for k in ('loss', 'nll', 'kl', 'smoothness', 'mask_sum'):
            assert abs(a[k] - b[k]) <= 2e-4 * max(abs(a[k]), 1e-3), (what, k, a[k], b[k])
        assert abs(a['lr'] - b['lr']) < 1e-12

    for i in range(7):                                     # 3 eager warm-up steps, the capture, 3 more replays
        close(dp_e.step(data, mask, prior), dp_g.step(data, mask, prior), 'step %d' % i)
    assert dp_g._g['graph'] is not None and dp_g.step_no == dp_e.step_no == 7
    assert dp_g.trainer._calls == dp_e.trainer._calls
    for a, b in zip(enc_e.parameters(), enc_g.parameters()):
        assert float((a - b).abs().max()) <= 2e-3 * float(a.abs().max()) + 1e-6
    # a replay is one graph launch: no kernel of this library is launched from the host during it
    before = qb.launch_count()
    torch.cuda.set_sync_debug_mode('error')
    try:
        s_g = dp_g.step(data, mask, prior)
    finally:
        torch.cuda.set_sync_debug_mode('default')
    assert qb.launch_count() == before
    close(dp_e.step(data, mask, prior), s_g, 'replay under sync-debug')
    # new batch contents go through the static buffers
    data2 = data * 1.01
    close(dp_e.step(data2, mask, prior), dp_g.step(data2, mask, prior), 'second batch')
    # checkpoint round trip into a fresh captured trainer: counters, schedule position and moments carry over
    path = str(tmp_path / 'dp.pt')
    dp_g.save(path)
    enc_r = copy.deepcopy(enc_g)
    dp_r = DataParallelTrainer(enc_r, _trainer(qb, cfg_noise_off, seed=11), layer, ft_lr=2e-3, cuda_graph=graph_mode)
    dp_r.load(path)
    for i in range(5):
        close(dp_e.step(data, mask, prior), dp_r.step(data, mask, prior), 'resumed step %d' % i)
    assert dp_r._g['graph'] is not None


# ---------------------------------------------------------------------------------- whole-volume inference (config 4)
def test_likelihood_map_and_posterior_inference(qb, dev, cfg_noise_off, physics):
    e = golden('ref_shim_elbo_optimal.npz')
    tr = _trainer(qb, cfg_noise_off, seed=5)
    layer = qb.SignalGenerationLayer(cfg_noise_off, True, True)
    n, S = e['q'].shape[0], 12
    eps = np.random.default_rng(4).standard_normal((n, S, 2)).astype(np.float32)
    got = tr.likelihood_map(layer, _t(e['q'], dev), _t(e['sigma'], dev), _t(e['data'], dev), _t(e['mask'], dev),
                            no_samples=S, eps=_t(eps, dev)).cpu().numpy().reshape(-1)
    yt = np.concatenate([e['data'], e['mask'][:, None]], -1)
    ref = np.zeros(n)
    for s in range(S):                                          # model.py:810-817, one forward pass per sample
        smp, _ = o.reparam_sample(e['q'], eps[:, s], True, np.float64)
        pred = o.forward(physics, smp, dtype=np.float64)
        ref += o.fine_tune_nll(yt, pred, e['sigma'], 2, np.float64, return_mean=False).reshape(-1)
    assert rel_max(got, ref / S) < GRAD_TOL
    # config 4 entry point: 64 samples per voxel on a volume, in-kernel Philox draws
    q = _t(e['q'], dev).reshape(2, 4, 4, 2, 5)
    res = tr.posterior_inference(layer, q, _t(e['sigma'], dev).reshape(2, 4, 4, 2, 11),
                                 _t(e['data'], dev).reshape(2, 4, 4, 2, 11), _t(e['mask'], dev).reshape(2, 4, 4, 2, 1),
                                 prior=_t(e['prior'], dev).reshape(2, 4, 4, 2, 5), no_samples=64)
    assert tuple(res['means'].shape) == (2, 4, 4, 2, 3) and tuple(res['variances'].shape) == (2, 4, 4, 2, 3)
    assert tuple(res['likelihood'].shape) == (2, 4, 4, 2, 1) and tuple(res['kl'].shape) == (2, 4, 4, 2, 1)
    for v in res.values():
        assert torch.isfinite(v).all()
    m = res['means'].reshape(-1, 3)
    assert float(m[:, 0].min()) > 0.04 and float(m[:, 0].max()) < 0.84 and float(res['variances'].min()) >= 0
    dead = _t(e['mask'], dev) == 0
    assert float(res['likelihood'].reshape(-1)[dead].abs().max()) == 0 and float(res['kl'].reshape(-1)[dead].abs().max()) == 0


# ---------------------------------------------------------------------------------- remaining rows of the scope table
def test_misalignment_augmentation(qb, dev, cfg_noise_off):
    """signals.py:80-96: images up to index 4 are never touched; a misaligned voxel's later images come from
    perturbed OEF/DBV; probability 1 hits every voxel, the observed fraction follows the probability."""
    x = _t(_rand_voxels(4096, 12), dev)
    clean = qb.SignalGenerationLayer(cfg_noise_off, True, True)(x)
    torch.manual_seed(3)
    aug = qb.SignalGenerationLayer(cfg_noise_off, True, True, misaligned_prob=1.0)(x)
    assert torch.equal(aug[:, :5], clean[:, :5])
    changed = (aug != clean).any(-1)
    assert float(changed.float().mean()) > 0.99
    first = (aug != clean).float().argmax(-1)[changed]
    assert int(first.min()) >= 5 and int(first.max()) <= 10              # from_index in [4, n_tau-1) -> first changed image 5..10
    torch.manual_seed(4)
    aug = qb.SignalGenerationLayer(cfg_noise_off, True, True, misaligned_prob=0.25)(x)
    frac = float((aug != clean).any(-1).float().mean())
    assert 0.2 < frac < 0.3 and torch.isfinite(aug).all()


def _marginals_from_labels(d):
    grid = np.empty((529, 2), np.float32)
    grid[d['perm']] = d['train_y'][:, :2]
    g = grid.reshape(23, 23, 2)
    return g[:, 0, 0].copy(), g[0, :, 1].copy()


def test_generation_with_misalignment_matches_reference_source(qb, dev, physics):
    """create_synthetic_dataset(..., misaligned_prob=0.3) of the reference source, its recorded draws replayed through
    generate_from_marginals -> qbold_generate + qbold_misalign + the chunked noise (signals.py:80-96 inside :282-285)."""
    d = golden('ref_shim_dataset_misalign.npz')
    cfg = o.default_config()                                                         # noise ON, as in the INI
    oefs, dbvs = _marginals_from_labels(d)
    layer = qb.SignalGenerationLayer(cfg, True, True, misaligned_prob=float(d['prob']))
    x, y = qb.generate_from_marginals(layer, _t(oefs, dev), _t(dbvs, dev),
                                      torch.as_tensor(d['perm'], device=dev).contiguous(), n_chunks=10,
                                      snr_u01=_t(d['snr_u01'].reshape(-1), dev), noise_eps=_t(d['noise_eps'], dev),
                                      misalign_u01=_t(d['mis_u01'], dev),
                                      misalign_index=torch.as_tensor(d['mis_index'], device=dev),
                                      misalign_eps=_t(d['mis_eps'], dev))
    assert tuple(x.shape) == (520, 11)
    assert rel_elem(y.cpu().numpy(), d['train_y']) < 1e-6
    assert rel_elem(x.cpu().numpy(), d['train_x']) < 2 * SIG_TOL
    # the production path: in-kernel Philox draws == oracle/philox.py draws, through the public entry point's layer
    ph = o.parse_params(dict(cfg, simulate_noise='False'))
    clean = qb.SignalGenerationLayer(dict(cfg, simulate_noise='False'), True, True, misaligned_prob=0.1, seed=5)
    xv = _rand_voxels(6000, 14)
    got = clean.misalign(_t(xv, dev), clean._forward_raw(_t(xv, dev)), seed=1234, offset=777)
    u, idx, eps = philox.misalign_draws(1234, np.arange(6000, dtype=np.uint64) + np.uint64(777), 11)
    want = o.forward_misaligned(ph, xv, 0.1, u, idx, eps, dtype=np.float64)
    assert rel_elem(got.cpu().numpy(), want) < SIG_TOL
    assert 0.07 < float((u < 0.1).mean()) < 0.13 and idx.min() == 4 and idx.max() == 9
    # create_synthetic_dataset honours misaligned_prob (the reference CLI passes 0.1, signals.py:330)
    cfg2 = dict(cfg, sample_size='100', simulate_noise='False')
    x0, y0 = qb.create_synthetic_dataset(cfg2, True, True, 0.0, device=dev, seed=3)
    x1, y1 = qb.create_synthetic_dataset(cfg2, True, True, 0.1, device=dev, seed=3)
    assert torch.equal(y0, y1) and torch.equal(x0[:, :5], x1[:, :5])
    frac = float((x0 != x1).any(-1).float().mean())
    assert 0.08 < frac < 0.12
    with pytest.raises(qb.QboldError):                                               # randint(4, 4) in the reference
        cfg5 = dict(cfg2, tau_start='0.0', tau_end='0.05', tau_step='0.01')
        qb.SignalGenerationLayer(cfg5, True, True, misaligned_prob=0.5)(_t(xv[:8], dev))


@pytest.mark.parametrize('full', [True, False])
@pytest.mark.parametrize('blood', [True, False])
def test_variable_hct_forward_and_gradient(qb, dev, cfg_noise_off, physics, full, blood):
    """variable_hct=True (signals.py:64-70): rows (OEF, DBV, Hct); value and the 3-column VJP (d/dHct through dw and
    the blood term) against the reference-source fixture and the float64 oracle; autograd through the layer."""
    g = golden('ref_shim_forward_hct.npz')
    key = 'f%d_b%d' % (int(full), int(blood))
    layer = qb.SignalGenerationLayer(cfg_noise_off, full, blood, variable_hct=True)
    x = _t(g['oef_dbv_hct'], dev)
    s, g2 = layer.forward_backward(x, _t(g['g_rand'], dev))
    _, g1 = layer.forward_backward(x, None, want_signal=False)
    assert tuple(g2.shape) == (96, 3)
    assert rel_elem(s.cpu().numpy(), g['signal_' + key]) < SIG_TOL
    for j in range(3):
        assert rel_max(g1.cpu().numpy()[:, j], g['grad_ones_' + key][:, j]) < GRAD_TOL
        assert rel_max(g2.cpu().numpy()[:, j], g['grad_rand_' + key][:, j]) < GRAD_TOL
    r = np.random.default_rng(17)
    xv = _rand_voxels(3000, 15)
    # Hct up to 0.6, but OEF * Hct <= 0.38: beyond 1.5 tau dw u_0 = 3.45e-4 the float32 reference's node 0 stops being
    # exactly dead (1 - j0f(x0) becomes one ulp = 6e-8, times a weight of 1.7e7) and float64 is no longer its oracle
    hct = np.minimum(r.uniform(0.2, 0.6, 3000), 0.38 / xv[:, 0]).astype(np.float32)
    xv = np.concatenate([xv, hct[:, None]], -1)
    gs = r.standard_normal((3000, 11)).astype(np.float32)
    s, gg = layer.forward_backward(_t(xv, dev), _t(gs, dev))
    s64, g64 = o.forward_backward(physics, xv, gs, full, blood, np.float64, variable_hct=True)
    assert rel_elem(s.cpu().numpy(), s64) < SIG_TOL
    for j in range(3):
        assert rel_max(gg.cpu().numpy()[:, j], g64[:, j]) < GRAD_TOL
    xa = _t(xv[:257], dev).reshape(257, 1, 3).requires_grad_(True)
    w = _t(gs[:257], dev).reshape(257, 1, 11)
    (layer(xa) * w).sum().backward()
    assert torch.equal(xa.grad.reshape(-1, 3), gg[:257])


def test_variable_hct_generation_and_misalignment(qb, dev, cfg_noise_off):
    d = golden('ref_shim_dataset_hct.npz')
    cfg = o.default_config()
    oefs, dbvs = _marginals_from_labels(d)
    layer = qb.SignalGenerationLayer(cfg, True, True, variable_hct=True)
    x, y = qb.generate_from_marginals(layer, _t(oefs, dev), _t(dbvs, dev),
                                      torch.as_tensor(d['perm'], device=dev).contiguous(), n_chunks=10,
                                      snr_u01=_t(d['snr_u01'].reshape(-1), dev), noise_eps=_t(d['noise_eps'], dev))
    assert tuple(x.shape) == (520, 11) and tuple(y.shape) == (529, 3)
    assert rel_elem(y.cpu().numpy(), d['train_y']) < 1e-6
    assert rel_elem(x.cpu().numpy(), d['train_x']) < 2 * SIG_TOL
    tx, ty = qb.create_synthetic_dataset(dict(cfg, sample_size='40'), True, True, 0.0, variable_hct=True, device=dev)
    assert tuple(tx.shape) == (1600, 11) and tuple(ty.shape) == (1600, 3) and torch.isfinite(tx).all()
    # misalignment with a per-voxel Hct column (signals.py:80-96 broadcasts hct [N,1] over the images)
    g = golden('ref_shim_forward_hct.npz')
    lay = qb.SignalGenerationLayer(cfg_noise_off, True, True, misaligned_prob=float(g['mis_prob']), variable_hct=True)
    xh = _t(g['oef_dbv_hct'], dev)
    got = lay.misalign(xh, lay._forward_raw(xh), _t(g['mis_u01'], dev), torch.as_tensor(g['mis_index'], device=dev),
                       _t(g['mis_eps'], dev))
    assert rel_elem(got.cpu().numpy(), g['signal_misaligned']) < SIG_TOL


def test_diagonal_posterior_variant(qb, dev, cfg_noise_off):
    """use_mvg=False (model.py:33-37, 695-708): 4 parameters per voxel, analytic KL == two 1-D Gaussian KLs."""
    e = golden('ref_shim_elbo_optimal.npz')
    tr = _trainer(qb, cfg_noise_off, use_mvg=False)
    q4, p4 = e['q'][:, :4], e['prior'][:, :4]
    smp = qb.ReparamTrickLayer(tr)((_t(q4, dev), None), eps=_t(e['eps'], dev)).cpu().numpy()
    sd_o, sd_d = np.exp(np.tanh(q4[:, 1]) * 3 - 1), np.exp(np.tanh(q4[:, 3]) * 3 - 1)
    z = np.stack([q4[:, 0] + e['eps'][:, 0] * sd_o, q4[:, 2] + e['eps'][:, 1] * sd_d], -1)
    assert rel_elem(smp, o.forward_transform(z, np.float64)) < SIG_TOL
    true = torch.cat([_t(p4, dev), _t(e['mask'], dev)[:, None]], -1)
    kl = tr.kl_loss(true, _t(q4, dev), return_mean=False).cpu().numpy().reshape(-1)

    def kl1(mq, lq, mp, lp):                                              # KL(N(mq, e^lq) || N(mp, e^lp))
        return lp - lq + (np.exp(2 * lq) + (mq - mp) ** 2) / (2 * np.exp(2 * lp)) - 0.5
    ls = lambda a: np.tanh(a.astype(np.float64)) * 3 - 1
    ref = kl1(q4[:, 0], ls(q4[:, 1]), p4[:, 0], ls(p4[:, 1])) + kl1(q4[:, 2], ls(q4[:, 3]), p4[:, 2], ls(p4[:, 3]))
    assert rel_max(kl, np.where(e['mask'] > 0, ref, 0)) < 1e-5


def test_fine_tuner_module_outputs(qb, dev, cfg_noise_off):
    from qbold_vi_b200.encoder import Encoder
    torch.manual_seed(0)
    tr = _trainer(qb, cfg_noise_off)
    layer = qb.SignalGenerationLayer(cfg_noise_off, True, True)
    ft = tr.build_fine_tuner(Encoder(no_units=8, no_intermediate_layers=1).to(dev), layer)
    data = torch.rand(1, 5, 4, 2, 11, device=dev) + 0.5
    mask = torch.ones(1, 5, 4, 2, 1, device=dev)
    out = ft(data, mask)
    assert tuple(out['predictions'].shape) == (1, 5, 4, 2, 5) and tuple(out['predicted_images'].shape) == (1, 5, 4, 2, 22)
    nll = tr.fine_tune_loss_fn(torch.cat([data, mask], -1), out['predicted_images'])
    nll.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in ft.encoder.parameters() if p.requires_grad)
    loss, info = ft.fused_loss(data, mask, out['predictions'].detach())
    assert torch.isfinite(loss)


# ---------------------------------------------------------------------------------- other parameter sets / tau grids
@pytest.mark.parametrize('case', ['irregular_taus', 'other_physics', 'five_taus', 'single_column'])
def test_other_parameter_sets(qb, dev, cfg_noise_off, case):
    """The scheduled quadrature is built from the ACTUAL tau array and physics; exercise grids other than optimal.yaml's."""
    cfg = dict(cfg_noise_off)
    taus = None
    if case == 'irregular_taus':                       # 7 taus, 6 distinct |tau| with irregular ratios
        taus = np.array([-0.011, -0.003, 0.0, 0.003, 0.017, 0.029, 0.061], np.float32)
    elif case == 'other_physics':
        # keeps 1.5*tau*dw*u_0 below 3.45e-4: beyond it node 0 of the float32 reference stops being exactly dead and
        # its value is quantisation noise of 1 - (1 - x^2/4) -- parity is ill-defined there
        cfg.update(b0='4.0', te='0.08', r2t='20.0', hct='0.42', dchi='2.0e-7')
    elif case == 'five_taus':
        cfg.update(tau_start='0.0', tau_end='0.05', tau_step='0.01')
    elif case == 'single_column':
        taus = np.array([0.0, 0.03], np.float32)
    ph = o.parse_params(cfg, taus=taus)
    layer = qb.SignalGenerationLayer(cfg, True, True, taus=taus)
    assert layer.params.sched_phases > 0
    nt = ph.n_tau
    x = _rand_voxels(1500, 31)
    gs = np.random.default_rng(2).standard_normal((1500, nt)).astype(np.float32)
    s, g = layer.forward_backward(_t(x, dev), _t(gs, dev))
    s64, g64 = o.forward_backward(ph, x, gs, dtype=np.float64)
    assert rel_elem(s.cpu().numpy(), s64) < SIG_TOL
    assert rel_max(g.cpu().numpy()[:, 0], g64[:, 0]) < GRAD_TOL and rel_max(g.cpu().numpy()[:, 1], g64[:, 1]) < GRAD_TOL
    assert rel_elem(layer(_t(x, dev)).cpu().numpy(), s64) < SIG_TOL              # forward-only kernel


def test_streaming_pretrainer(qb, dev):
    """train.py:379-427 with on-the-fly generation: the pre-training NLL falls and the OEF error shrinks."""
    from qbold_vi_b200.encoder import Encoder
    from qbold_vi_b200.distributed import StreamingPretrainer
    cfg = o.default_config()
    cfg['sample_size'] = '400'
    torch.manual_seed(2)
    enc = Encoder().to(dev)
    tr = _trainer(qb, cfg)
    pt = StreamingPretrainer(enc, tr, cfg, uniform_prop=0.1, lr=2e-3, batch_blocks=16, seed=5, device=dev)
    x, y = pt.next_batch()
    assert tuple(x.shape) == (8000, 11) and tuple(y.shape) == (8000, 3) and torch.isfinite(x).all()
    assert float(y[:, 0].min()) >= 0.05 - 1e-6 and float(y[:, 1].max()) <= 0.195 + 1e-6
    stats = [pt.step() for _ in range(40)]
    assert all(np.isfinite(s['loss']) for s in stats)
    assert np.mean([s['loss'] for s in stats[-5:]]) < np.mean([s['loss'] for s in stats[:5]]) - 0.5
    assert np.mean([s['oef_mse'] for s in stats[-5:]]) < np.mean([s['oef_mse'] for s in stats[:5]])


@pytest.mark.parametrize('grid', ['tau24', 'tau5', 'tau11_closed_form'])
def test_fused_elbo_other_grids_and_closed_form(qb, dev, cfg_noise_off, grid):
    """The unpaired / column-major ELBO kernels (24 taus: 16 columns, n_tau > 16), a 5-tau grid on the paired
    kernel, and the closed-form KL inside the fused kernel, all against the float64 oracle."""
    cfg = dict(cfg_noise_off)
    S = 70
    if grid == 'tau24':
        cfg.update(tau_start='-0.028', tau_end='0.065', tau_step='0.004')
    elif grid == 'tau5':
        cfg.update(tau_start='-0.01', tau_end='0.035', tau_step='0.01')
    ph = o.parse_params(cfg)
    nt = ph.n_tau
    se = int(abs(float(cfg['tau_start']) / float(cfg['tau_step'])))
    n = 300
    r = np.random.default_rng(33)
    q = np.stack([r.normal(-0.3, 0.7, n), r.normal(0, 0.5, n), r.normal(-1.2, 0.7, n), r.normal(0, 0.5, n),
                  r.normal(0, 0.8, n)], -1).astype(np.float32)
    prior = (q + r.normal(0, 0.3, (n, 5))).astype(np.float32)
    sigma = np.exp(r.normal(np.log(0.05), 0.2, (n, nt))).astype(np.float32)
    mask = (r.uniform(size=n) > 0.3).astype(np.float32)
    truth = np.stack([r.uniform(0.1, 0.7, n), r.uniform(0.005, 0.15, n)], -1)
    data = (o.forward(ph, truth, dtype=np.float64) * 100 * (1 + 0.02 * r.standard_normal((n, nt)))).astype(np.float32)
    data *= mask[:, None]
    eps = r.standard_normal((n, 2)).astype(np.float32)
    eps_kl = r.standard_normal((n, S, 2)).astype(np.float32)
    tr = _trainer(qb, cfg)
    assert tr._se_idx == se
    layer = qb.SignalGenerationLayer(cfg, True, True)
    qt, st = _t(q, dev).requires_grad_(True), _t(sigma, dev).requires_grad_(True)
    if grid == 'tau11_closed_form':
        ref = o.elbo_and_grads(ph, q, sigma, data, mask, prior, eps, None, np.float64, se_idx=se)
        klc = np.where(mask > 0, o.closed_form_kl(prior, q), 0)
        loss, info = tr.fused_elbo(layer, qt, st, _t(data, dev), _t(mask, dev), _t(prior, dev), kl_samples=0,
                                   eps=_t(eps, dev), return_maps=True)
        assert rel_max(info['kl_map'].cpu().numpy(), klc) < 1e-5
        assert rel_elem(info['kl'].item(), klc.sum() / mask.sum()) < 1e-5
        assert rel_elem(info['nll'].item(), ref['nll']) < GRAD_TOL
        loss.backward()
        # gradient of the closed form by float64 central differences, added to the oracle's likelihood gradient
        g_kl = np.zeros((n, 5))
        for j in range(5):
            d = np.zeros(5)
            d[j] = 1e-6
            g_kl[:, j] = (o.closed_form_kl(prior, q.astype(np.float64) + d) - o.closed_form_kl(prior, q.astype(np.float64) - d)) / 2e-6
        g_ref = ref['grad_q'] + np.where(mask[:, None] > 0, g_kl, 0) / mask.sum()
        assert rel_max(qt.grad.cpu().numpy(), g_ref) < GRAD_TOL
        return
    ref = o.elbo_and_grads(ph, q, sigma, data, mask, prior, eps, eps_kl, np.float64, se_idx=se)
    loss, info = tr.fused_elbo(layer, qt, st, _t(data, dev), _t(mask, dev), _t(prior, dev), kl_samples=S,
                               eps=_t(eps, dev), eps_kl=_t(eps_kl, dev), return_maps=True)
    loss.backward()
    assert rel_elem(info['nll'].item(), ref['nll']) < GRAD_TOL and rel_elem(info['kl'].item(), ref['kl']) < GRAD_TOL
    assert rel_max(info['nll_map'].cpu().numpy(), ref['nll_map']) < GRAD_TOL
    assert rel_max(qt.grad.cpu().numpy(), ref['grad_q']) < GRAD_TOL
    assert rel_max(st.grad.cpu().numpy(), ref['grad_sigma']) < GRAD_TOL


@pytest.mark.parametrize('tag', ['optimal', 'multinorm', 'studentt'])
def test_standalone_fine_tune_loss_fn_kernel(qb, dev, cfg_noise_off, tag):
    """fine_tune_loss_fn(y_true, y_pred) on predictions that already exist (qbold_nll) vs the reference source."""
    e = golden('ref_shim_elbo_%s.npz' % tag)
    tr = _trainer(qb, cfg_noise_off, student_t_df=float(e['student_t_df']),
                  multi_image_normalisation=bool(e['multi_image_normalisation']))
    y_true = torch.cat([_t(e['data'], dev), _t(e['mask'], dev)[:, None]], -1).reshape(2, 4, 4, 2, 12)
    pred = _t(e['pred'], dev).requires_grad_(True)
    sg = _t(e['sigma'], dev).requires_grad_(True)
    y_pred = torch.cat([pred, sg], -1).reshape(2, 4, 4, 2, 22)
    nll = tr.fine_tune_loss_fn(y_true, y_pred)
    assert rel_elem(nll.item(), e['nll']) < GRAD_TOL
    nll.backward()
    assert rel_max(sg.grad.cpu().numpy(), e['grad_sigma']) < GRAD_TOL
    m = tr.fine_tune_loss_fn(y_true, y_pred, return_mean=False)
    assert tuple(m.shape) == (64, 1) and rel_max(m.detach().cpu().numpy().reshape(-1), e['nll_map']) < GRAD_TOL
    # d nll / d pred against autograd of the reference formulas written with torch ops (float64)
    p64 = _t(e['pred'], dev).double().requires_grad_(True)
    yt, s64, mk = _t(e['data'], dev).double(), _t(e['sigma'], dev).double(), _t(e['mask'], dev).double()
    se = 2
    if bool(e['multi_image_normalisation']):
        yn, pn = yt / (yt[:, se - 1:se + 2].mean(-1, keepdim=True) + 1e-3), p64 / (p64[:, se - 1:se + 2].mean(-1, keepdim=True) + 1e-3)
    else:
        yn, pn = yt / (yt[:, se:se + 1] + 1e-3), p64 / (p64[:, se:se + 1] + 1e-3)
    zq = (yn - pn) / s64
    df = float(e['student_t_df'])
    if df < 50:
        import math
        c = math.lgamma(0.5 * (df + 1)) - math.lgamma(0.5 * df) - 0.5 * math.log(df * math.pi)
        nl = -(c - torch.log(s64) - 0.5 * (df + 1) * torch.log1p(zq * zq / df))
    else:
        nl = torch.log(s64) + 0.5 * math_log_2pi() + 0.5 * zq * zq
    ref = ((nl.sum(-1) * mk).sum() / mk.sum())
    ref.backward()
    assert rel_max(pred.grad.cpu().numpy(), p64.grad.cpu().numpy()) < GRAD_TOL


def math_log_2pi():
    import math
    return math.log(2.0 * math.pi)


# ---------------------------------------------------------------- losses either side of the path
def _adj():
    a = golden('ref_shim_adjacent.npz')
    return a, tuple(int(v) for v in a['shape'])


@pytest.mark.parametrize('tag,c', [('mvg', 5), ('diag', 4)])
def test_smoothness_kernel_matches_reference_source(qb, dev, cfg_noise_off, tag, c):
    a, shp = _adj()
    tr = _trainer(qb, cfg_noise_off, use_mvg=(c == 5))
    q = _t(a['q5'][:, :c].reshape(shp + (c,)), dev).requires_grad_(True)
    true = torch.cat([_t(a['prior5'][:, :c].reshape(shp + (c,)), dev), _t(a['mask'].reshape(shp + (1,)), dev)], -1)
    tv = tr.smoothness_loss(true, q)
    assert rel_elem(tv.item(), a['tv_' + tag]) < GRAD_TOL
    (3.0 * tv).backward()
    assert rel_max(q.grad.cpu().numpy().reshape(-1, c) / 3.0, a['tv_%s_grad' % tag]) < GRAD_TOL
    val, g = o.smoothness_loss(a['q5'][:, :c].reshape(shp + (c,)), a['mask'].reshape(shp))
    assert rel_elem(tv.item(), val) < 1e-5


def test_smoothness_kernel_large_volume_properties(qb, dev, cfg_noise_off):
    """64^3 volumes: against the same formula in float64 torch ops, and translation invariance of the value."""
    tr = _trainer(qb, cfg_noise_off)
    g = torch.Generator(device='cpu').manual_seed(3)
    q = (torch.randn((2, 64, 64, 64, 5), generator=g) * 0.8).to(dev).requires_grad_(True)
    mask = (torch.rand((2, 64, 64, 64, 1), generator=g) > 0.2).float().to(dev)
    true = torch.cat([torch.zeros_like(q.detach()), mask], -1)
    tv = tr.smoothness_loss(true, q)
    tv.backward()
    q64 = q.detach().double().requires_grad_(True)
    p = torch.sigmoid(torch.stack([q64[..., 0], q64[..., 2]], -1))
    p = (p * torch.tensor([0.8, 0.2], device=dev, dtype=torch.float64) +
         torch.tensor([0.04, 0.001], device=dev, dtype=torch.float64)) / torch.tensor([0.8, 0.2], device=dev,
                                                                                        dtype=torch.float64)
    m = mask.double()
    ref = (((p[:, :-1] - p[:, 1:]).abs() * (m[:, :-1] * m[:, 1:])).sum() +
           ((p[:, :, :-1] - p[:, :, 1:]).abs() * (m[:, :, :-1] * m[:, :, 1:])).sum()) / m.sum()
    ref.backward()
    assert rel_elem(tv.item(), ref.item()) < 1e-5
    # sign flips of |d| for float32-vs-float64 near-ties are the only legitimate difference: allow a few voxels
    diff = (q.grad.double() - q64.grad).abs().amax(-1) > 1e-4 * q64.grad.abs().max()
    assert diff.float().mean().item() < 1e-4
    assert float(q.grad[..., [1, 3, 4]].abs().max()) == 0.0
    rolled = tr.smoothness_loss(torch.roll(true, 1, 3), torch.roll(q.detach(), 1, 3))     # z has no neighbours
    assert rel_elem(rolled.item(), tv.item()) < 1e-6


@pytest.mark.parametrize('tag,use_mvg,ig', [('mvg', True, (0.0, 0.0)), ('mvg_ig', True, (3.0, 0.15)),
                                            ('diag', False, (0.0, 0.0)), ('diag_ig', False, (3.0, 0.15))])
def test_synthetic_data_loss_kernel_matches_reference_source(qb, dev, cfg_noise_off, tag, use_mvg, ig):
    a, shp = _adj()
    c = 5 if use_mvg else 4
    tr = _trainer(qb, cfg_noise_off, use_mvg=use_mvg)
    pred = _t(a['q5'][:, :c].reshape(shp + (c,)), dev).requires_grad_(True)
    labels = _t(a['synth_%s_labels' % tag].reshape(shp + (3,)), dev)
    loss = tr.synthetic_data_loss(labels, pred, False, ig[0], ig[1])
    assert rel_elem(loss.item(), a['synth_' + tag]) < GRAD_TOL
    loss.backward()
    assert rel_max(pred.grad.cpu().numpy().reshape(-1, c), a['synth_%s_grad' % tag]) < GRAD_TOL
    rows, g = o.synthetic_data_nll(a['synth_%s_labels' % tag], a['q5'][:, :c], use_mvg, ig[0], ig[1], dt=np.float32)
    assert rel_max(pred.grad.cpu().numpy().reshape(-1, c), g / rows.shape[0]) < 1e-5


def test_synthetic_data_loss_r2p_term_matches_reference_source(qb, dev, cfg_noise_off):
    a, shp = _adj()
    tr = _trainer(qb, cfg_noise_off)
    pred = _t(a['q5'].reshape(shp + (5,)), dev).requires_grad_(True)
    loss = tr.synthetic_data_loss(_t(a['labels'].reshape(shp + (3,)), dev), pred, True, 0.0, 0.0,
                                  eps=_t(a['synth_r2p_eps'], dev))
    assert rel_elem(loss.item(), a['synth_r2p']) < GRAD_TOL
    loss.backward()
    assert rel_max(pred.grad.cpu().numpy().reshape(-1, 5), a['synth_r2p_grad']) < GRAD_TOL


@pytest.mark.parametrize('i,name', [(0, 'oef'), (1, 'dbv'), (2, 'r2p')])
def test_pretraining_metrics_match_reference_source(qb, dev, cfg_noise_off, i, name):
    a, shp = _adj()
    tr = _trainer(qb, cfg_noise_off)
    val = getattr(tr, name + '_metric')(_t(a['labels'].reshape(shp + (3,)), dev), _t(a['q5'].reshape(shp + (5,)), dev),
                                        eps=_t(a['metric_%s_eps' % name], dev))
    assert rel_elem(val.item(), a['metric_' + name]) < GRAD_TOL


def test_diagonal_kl_variants_match_reference_source(qb, dev, cfg_noise_off):
    a, shp = _adj()
    mask = _t(a['mask'].reshape(shp + (1,)), dev)
    true = torch.cat([_t(a['prior5'][:, :4].reshape(shp + (4,)), dev), mask], -1)
    tr = _trainer(qb, cfg_noise_off, use_mvg=False)
    q = _t(a['q5'][:, :4].reshape(shp + (4,)), dev).requires_grad_(True)
    kl = tr.kl_loss(true, q)
    assert rel_elem(kl.item(), a['kl_diag']) < GRAD_TOL
    kl.backward()
    assert rel_max(q.grad.cpu().numpy().reshape(-1, 4), a['kl_diag_grad']) < GRAD_TOL
    kmap = tr.kl_loss(true, q.detach(), return_mean=False)
    ref, _, _ = o.diag_kl(a['prior5'][:, :4], a['q5'][:, :4])
    assert tuple(kmap.shape) == shp + (1,)
    assert rel_max(kmap.cpu().numpy().reshape(-1), ref * (a['mask'] > 0)) < 1e-5
    # population prior: q and the trainable prior side by side, InverseGamma(1,2) hyper-prior on its mean log-std
    trp = _trainer(qb, cfg_noise_off, use_mvg=False, use_population_prior=True)
    q8 = _t(a['kl_pop_pred'].reshape(shp + (8,)), dev).requires_grad_(True)
    klp = trp.kl_loss(true, q8)
    assert rel_elem(klp.item(), a['kl_pop']) < GRAD_TOL
    klp.backward()
    assert rel_max(q8.grad.cpu().numpy().reshape(-1, 8), a['kl_pop_grad']) < GRAD_TOL


def test_save_predictions_writes_posterior_maps(qb, dev, cfg_noise_off, tmp_path):
    """save_predictions (model.py:772-887): NIfTI maps of means / variances / likelihood / KL / residual."""
    from qbold_vi_b200.encoder import Encoder
    from qbold_vi_b200.nifti import load_nifti
    torch.manual_seed(0)
    tr = _trainer(qb, cfg_noise_off)
    layer = qb.SignalGenerationLayer(cfg_noise_off, True, True)
    enc = Encoder(no_units=8, no_intermediate_layers=1).to(dev)
    ft = tr.build_fine_tuner(enc, layer)
    g = torch.Generator().manual_seed(4)
    truth = torch.stack([torch.rand((2, 6, 5, 3), generator=g) * 0.5 + 0.15,
                         torch.rand((2, 6, 5, 3), generator=g) * 0.1 + 0.01], -1).to(dev)
    images = layer(truth) * 100.0
    mask = (torch.rand((2, 6, 5, 3, 1), generator=g) > 0.2).float().to(dev)
    data = torch.cat([images, mask], -1).cpu().numpy()
    with torch.no_grad():
        priors = enc(images * mask)[0]
    base = str(tmp_path / 'subj')
    out = tr.save_predictions(enc, data, base, use_first_op=False, fine_tuner_model=ft, priors=priors)
    for name, last in (('_oef', 2), ('_dbv', 2), ('_r2p', 2), ('_logstds', 6), ('_likelihood', 2), ('_kl', 2),
                       ('_residual', 2)):
        arr, _ = load_nifti(base + name + '.nii.gz')
        assert arr.shape == (6, 5, 3, last), name
        assert np.isfinite(arr).all(), name
    oef, _ = load_nifti(base + '_oef.nii.gz')
    assert np.array_equal(oef[..., 1], out['means'][1, ..., 0])
    assert 0.04 <= oef.min() and oef.max() <= 0.84
    lik, _ = load_nifti(base + '_likelihood.nii.gz')
    assert np.all(lik[..., 0][mask[0, ..., 0].cpu().numpy() == 0] == 0.0)         # masked voxels carry no likelihood
    with pytest.raises(NotImplementedError):
        tr.save_predictions(enc, data, base, transform_directory='/nonexistent')


# ---------------------------------------------------------------------------------- robustness of the C ABI
def test_c_abi_argument_errors_and_messages(qb, dev, cfg_noise_off):
    """Every entry point rejects bad arguments with QBOLD_EINVAL (-1) and a message, and never launches."""
    import ctypes as C
    lib = qb._lib.lib()
    layer = qb.SignalGenerationLayer(cfg_noise_off, True, True)
    P = C.byref(layer.params)
    x = torch.rand((8, 2), device=dev)
    out = torch.empty((8, 11), device=dev)
    before = qb.launch_count()
    assert lib.qbold_forward(P, None, 2, 8, out.data_ptr(), None) == -1
    assert b'qbold_forward' in lib.qbold_last_error()
    assert lib.qbold_forward(P, x.data_ptr(), 5, 8, out.data_ptr(), None) == -1            # 2 or 3 channels only
    assert lib.qbold_forward(P, x.data_ptr(), 2, -1, out.data_ptr(), None) == -1
    assert lib.qbold_forward(None, x.data_ptr(), 2, 8, out.data_ptr(), None) == -1
    assert lib.qbold_forward_backward(P, x.data_ptr(), None, 8, None, None, None) == -1
    assert lib.qbold_kl(None, None, None, None, 0, 0, 70, 8, None, None, None) == -1
    assert lib.qbold_smoothness(x.data_ptr(), 3, x.data_ptr(), 1, 2, 2, 2, 1.0, None, out.data_ptr(), None) == -1
    assert lib.qbold_synth_nll(x.data_ptr(), 1, x.data_ptr(), 1, 0.0, 0.0, 8, 1.0, out.data_ptr(), None, None, None) == -1
    assert lib.qbold_diag_kl(x.data_ptr(), 2, x.data_ptr(), 4, None, 8, out.data_ptr(), None, 0, None, 0, None) == -1
    bad = qb._lib.QboldParams()
    C.memmove(C.byref(bad), P, C.sizeof(bad))
    bad.abi_version = 999
    assert lib.qbold_forward(C.byref(bad), x.data_ptr(), 2, 8, out.data_ptr(), None) == -1
    assert qb.launch_count() == before
    assert lib.qbold_forward(P, x.data_ptr(), 2, 0, out.data_ptr(), None) == 0              # empty is fine
    with pytest.raises(qb.QboldError):
        layer(torch.rand(4, 2))                                                             # CPU tensor: no CPU path


def test_non_default_stream_and_nan_propagation(qb, dev, cfg_noise_off, physics):
    layer = qb.SignalGenerationLayer(cfg_noise_off, True, True)
    x = _rand_voxels(4097, 5)
    ref = layer(_t(x, dev))
    s = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(s):
        xt = _t(x, dev)
        got = layer(xt)
        sig, grad = layer.forward_backward(xt, torch.ones((4097, 11), device=dev))
    s.synchronize()
    assert torch.equal(got, ref) and torch.equal(sig, ref)
    # NaN / Inf inputs propagate to that voxel only (the reference relies on TerminateOnNaN)
    xb = x.copy()
    xb[10] = [np.nan, 0.05]
    xb[20] = [0.4, np.inf]
    out = layer(_t(xb, dev)).cpu().numpy()
    assert np.isnan(out[10]).all() and not np.isfinite(out[20]).all()
    keep = np.ones(4097, bool)
    keep[[10, 20]] = False
    assert np.array_equal(out[keep], ref.cpu().numpy()[keep])


def test_more_than_2_31_output_elements(qb, dev, cfg_noise_off):
    """64-bit indexing: 200 M voxels x 11 taus = 2.2e9 output elements (8.8 GB) in one launch."""
    layer = qb.SignalGenerationLayer(cfg_noise_off, True, True)
    n = 200_000_000
    base = _t(_rand_voxels(1 << 20, 8), dev)
    x = base.repeat((n + (1 << 20) - 1) // (1 << 20), 1)[:n].contiguous()
    out = layer(x)
    assert tuple(out.shape) == (n, 11)
    ref = layer(base)
    tail = n - (n // (1 << 20)) * (1 << 20)
    assert torch.equal(out[-tail:], ref[:tail])                                             # beyond element 2^31
    assert torch.equal(out[(1 << 20) * 100:(1 << 20) * 101], ref)
    del out, x
    torch.cuda.empty_cache()


# ---------------------------------------------------------------------------------- stream-1 MLP on tcgen05
@pytest.mark.parametrize('n_tau,units,blocks,multi,n', [(11, 60, 2, False, 5000), (11, 60, 2, True, 128),
                                                       (24, 64, 1, False, 333), (11, 32, 3, False, 70000),
                                                       (5, 10, 1, False, 1)])
def test_voxelwise_encoder_tensor_core_kernel(qb, dev, n_tau, units, blocks, multi, n):
    """qbold_encoder_mlp_forward (TF32 tcgen05, fp32 accumulate) vs the same layers in float32 torch (TF32 off)."""
    from qbold_vi_b200.encoder import Encoder
    torch.manual_seed(n_tau * 100 + units)
    se = 2
    enc = Encoder(no_units=units, no_intermediate_layers=blocks, no_ip_images=n_tau, se_idx=se,
                  multi_image_normalisation=multi).to(dev)
    with torch.no_grad():                                                    # non-trivial biases
        for m in [enc.first, enc.final] + [b.pointwise for b in enc.blocks]:
            m.bias.normal_(0.0, 0.3)
    g = torch.Generator().manual_seed(n)
    data = (torch.rand((n, 1, 1, 1, n_tau), generator=g) * 150.0 + 20.0).to(dev)
    data[0, ..., 0] = 0.0                                                    # exercises the 1e-2 clip
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            ref = enc(data)[0]
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    got = enc.voxelwise_fused(data)
    assert got.shape == ref.shape
    # TF32 operands (10-bit mantissa), 3-4 chained layers: a few 1e-3 of the output scale
    scale = float(ref.abs().max())
    assert float((got - ref).abs().max()) < 1e-2 * scale, (float((got - ref).abs().max()), scale)
    assert float((got - ref).abs().mean()) < 2e-3 * scale


def test_fine_tune_dataset_prior_uses_the_tensor_core_branch(qb, dev):
    """FineTuneDataset (train.py:17-72): the prior is output 0 of the pre-trained model; on CUDA it comes from the
    one-launch tcgen05 branch and must agree with the full torch forward."""
    from qbold_vi_b200.data import FineTuneDataset
    from qbold_vi_b200.encoder import Encoder
    torch.manual_seed(2)
    enc = Encoder(no_units=60, no_intermediate_layers=2).to(dev)
    assert enc.supports_voxelwise_fused()
    real = torch.rand(2, 60, 50, 4, 12) * 100.0 + 20.0
    real[..., -1] = (real[..., -1] > 60.0).float()
    ds = FineTuneDataset(real, enc, crop_size=20, training=False, device=dev)
    masked = ds.real[..., :-1] * ds.real[..., -1:]
    with torch.no_grad():
        ref = enc(masked)[0]
    assert float((ds.prior - ref).abs().max()) < 1e-2 * float(ref.abs().max())
    (data, mask), tgt = next(iter(ds))
    assert tuple(data.shape) == (3, 20, 20, 4, 11) and tuple(tgt['predictions'].shape) == (3, 20, 20, 4, 6)
    gelu = Encoder(no_units=60, no_intermediate_layers=2, activation='gelu').to(dev)
    assert not gelu.supports_voxelwise_fused()
    with pytest.raises(qb.QboldError):
        gelu.voxelwise_fused(masked)


def test_mog_population_prior_kl_matches_reference_source(qb, dev, cfg_noise_off):
    a, shp = _adj()
    mask = _t(a['mask'].reshape(shp + (1,)), dev)
    true = torch.cat([_t(a['prior5'][:, :4].reshape(shp + (4,)), dev), mask], -1)
    tr = _trainer(qb, cfg_noise_off, use_mvg=False, use_population_prior=True, mog_components=3)
    q16 = _t(a['kl_mog_pred'].reshape(shp + (16,)), dev).requires_grad_(True)
    kl = tr.kl_loss(true, q16, eps=_t(a['kl_mog_eps'].reshape(shp + (2,)), dev))
    assert rel_elem(kl.item(), a['kl_mog']) < GRAD_TOL
    kl.backward()
    assert rel_max(q16.grad.cpu().numpy().reshape(-1, 16), a['kl_mog_grad']) < GRAD_TOL


def test_infer_inv_gamma_pretraining_loss_matches_reference_source(qb, dev, cfg_noise_off):
    """synthetic_data_loss with infer_inv_gamma=True (model.py:454-455,493-496): the learned InverseGamma parameters
    ride in channels 4..7; value and the gradient of all 8 channels (the hyper-parameter gradient lands on voxel 0)."""
    a, _ = _adj()
    tr = _trainer(qb, cfg_noise_off, use_mvg=False, infer_inv_gamma=True)
    q8 = _t(a['synth_iginf_pred'], dev).reshape(-1, 1, 1, 1, 8).requires_grad_(True)
    loss = tr.synthetic_data_loss(_t(a['synth_iginf_labels'], dev).reshape(-1, 1, 1, 1, 3), q8)
    loss.backward()
    assert rel_elem(loss.item(), a['synth_iginf']) < GRAD_TOL
    g = q8.grad.cpu().numpy().reshape(-1, 8)
    assert rel_max(g[:, :4], a['synth_iginf_grad'][:, :4]) < GRAD_TOL
    assert rel_elem(g[0, 4:], a['synth_iginf_grad'][0, 4:]) < GRAD_TOL and np.all(g[1:, 4:] == 0)
    with pytest.raises(ValueError):                                                  # tf.split(y_pred, 2, -1) of 9 channels
        _trainer(qb, cfg_noise_off, use_mvg=True, infer_inv_gamma=True).synthetic_data_loss(
            _t(a['synth_iginf_labels'], dev), torch.zeros(a['labels'].shape[0], 9, device=dev))
    # the encoder carries the hyper-prior variable (model.py:201-205)
    from qbold_vi_b200.encoder import Encoder
    enc = Encoder(no_units=16, no_intermediate_layers=1, use_mvg=False, infer_inv_gamma=True).to(dev)
    out0, out1, _ = enc(torch.rand(1, 4, 4, 2, 11, device=dev) + 0.5)
    assert out0.shape[-1] == 8 and out1.shape[-1] == 4
    assert torch.allclose(out0[0, 0, 0, 0, 4:], torch.tensor([20.0, 2.5, 20.0, 2.5], device=dev))


def test_population_prior_is_trainable_and_operands_are_validated(qb, dev, cfg_noise_off):
    from qbold_vi_b200.encoder import Encoder
    torch.manual_seed(0)
    layer = qb.SignalGenerationLayer(cfg_noise_off, True, True)
    data = layer(_t(_rand_voxels(2 * 6 * 6 * 2, 4), dev)).reshape(2, 6, 6, 2, 11) * 100.0
    mask = (torch.rand(2, 6, 6, 2, 1, device=dev) > 0.2).float()
    for m_comp in (1, 3):
        tr = _trainer(qb, cfg_noise_off, use_mvg=False, use_population_prior=True, mog_components=m_comp)
        enc = Encoder(no_units=16, no_intermediate_layers=1, use_mvg=False).to(dev)
        ft = tr.build_fine_tuner(enc, layer)
        assert ft.pop_prior.numel() == 4 * m_comp and ft.pop_prior.requires_grad
        out = ft(data * mask, mask)
        assert out['predictions'].shape[-1] == 4 * (m_comp + 1)                      # model.py:268-271
        loss, info = ft.fused_loss(data * mask, mask, None)
        loss.backward()
        assert torch.isfinite(loss) and ft.pop_prior.grad is not None and float(ft.pop_prior.grad.abs().sum()) > 0
        assert enc.final.weight.grad is not None and float(enc.final.weight.grad.abs().sum()) > 0
    with pytest.raises(NotImplementedError):
        _trainer(qb, cfg_noise_off, use_mvg=True, use_population_prior=True).build_fine_tuner(enc, layer)
    tr = _trainer(qb, cfg_noise_off)
    q10 = torch.zeros(8, 10, device=dev)
    true = torch.zeros(8, 6, device=dev)
    with pytest.raises(ValueError):
        tr.kl_loss(true, q10)                                                        # 10-channel 'predictions'
    with pytest.raises(ValueError):
        tr.fused_elbo(layer, torch.zeros(8, 5, device=dev), torch.ones(8, 11, device=dev), torch.ones(8, 11, device=dev),
                      torch.ones(8, 1, device=dev), torch.zeros(4, 5, device=dev))   # prior of another voxel count
    # the parameter block follows the trainer's flags and the layer's lifetime (no id()-keyed cache)
    tr._student_t_df = 4.0
    p1 = tr._params_for(layer)
    tr._student_t_df = 200
    p2 = tr._params_for(layer)
    assert p1 is not p2 and float(p1.student_t_df) == 4.0 and tr._params_for(layer) is p2


def test_fused_elbo_full_size_properties(qb, dev, cfg_noise_off):
    """4 M voxels through the fused kernel with in-kernel Philox: exact zeros on masked voxels, bit-exact repeat,
    bit-exact shard invariance (the counter is the global voxel index), map / sum consistency."""
    n = 1 << 22
    g = torch.Generator(device=dev).manual_seed(5)
    layer = qb.SignalGenerationLayer(cfg_noise_off, True, True)
    tr = _trainer(qb, cfg_noise_off)
    x = torch.rand((n, 2), device=dev, generator=g)
    x[:, 0] = x[:, 0] * 0.7 + 0.08
    x[:, 1] = x[:, 1] * 0.15 + 0.01
    q = torch.stack([torch.logit((x[:, 0] - 0.04) / 0.8), torch.randn(n, device=dev, generator=g) * 0.3,
                     torch.logit((x[:, 1] - 0.001) / 0.2), torch.randn(n, device=dev, generator=g) * 0.3,
                     torch.randn(n, device=dev, generator=g) * 0.5], -1).contiguous()
    prior = (q + 0.3 * torch.randn((n, 5), device=dev, generator=g)).contiguous()
    sigma = torch.exp(torch.randn((n, 11), device=dev, generator=g) * 0.2 - 3.0)
    mask = (torch.rand(n, device=dev, generator=g) > 0.4).float()
    data = layer(x) * 100.0 * mask[:, None]
    msum = float(mask.sum())

    def run(sl=slice(None), offset=0):
        qq = q[sl].clone().requires_grad_(True)
        sg = sigma[sl].clone().requires_grad_(True)
        loss, info = tr.fused_elbo(layer, qq, sg, data[sl], mask[sl], prior[sl], kl_samples=70, mask_sum=msum, seed=99,
                                   return_maps=True, offset=offset)
        loss.backward()
        return loss.detach(), info, qq.grad, sg.grad

    loss, info, gq, gs = run()
    assert torch.isfinite(loss) and float(info['non_finite']) == 0 and float(info['mask_sum']) == msum
    off = mask == 0
    assert float(gq[off].abs().max()) == 0.0 and float(gs[off].abs().max()) == 0.0
    assert float(info['nll_map'][off].abs().max()) == 0.0 and float(info['kl_map'][off].abs().max()) == 0.0
    assert rel_elem(float(info['nll_map'].double().sum() / msum), float(info['nll'])) < 1e-5
    assert rel_elem(float(info['kl_map'].double().sum() / msum), float(info['kl'])) < 1e-5
    _, info2, gq2, gs2 = run()
    assert torch.equal(gq, gq2) and torch.equal(gs, gs2) and torch.equal(info['kl_map'], info2['kl_map'])
    lo = 1_500_001                                                                    # odd start: pairs re-align
    sl = slice(lo, lo + 100_000)
    _, info3, gq3, gs3 = run(sl, offset=lo)
    assert torch.equal(gq3, gq[sl]) and torch.equal(gs3, gs[sl])
    assert torch.equal(info3['kl_map'], info['kl_map'][sl]) and torch.equal(info3['nll_map'], info['nll_map'][sl])
    # the KL estimator is unbiased for the closed form: per-voxel difference has zero mean within sampling error
    _, info0, _, _ = (lambda: (None, tr.fused_elbo(layer, q, sigma, data, mask, prior, kl_samples=0, mask_sum=msum,
                                                  return_maps=True)[1], None, None))()
    d = (info['kl_map'] - info0['kl_map'])[mask > 0].double()
    assert abs(float(d.mean())) < 5.0 * float(d.std()) / (d.numel() ** 0.5) + 1e-4 * float(info0['kl_map'].mean())


@pytest.mark.parametrize('relu', [False, True])
@pytest.mark.parametrize('n,n_in,n_out', [(5000, 60, 60), (70001, 56, 12), (129, 8, 64), (1, 4, 4)])
def test_dense_tensor_core_forward_and_input_gradient(qb, dev, monkeypatch, n, n_in, n_out, relu):
    """qbold_dense_tc (tcgen05 TF32): y = act(x W^T + b) and dx = (g * relu') W vs float64."""
    from qbold_vi_b200 import encoder as E
    from qbold_vi_b200.encoder import _DenseFn, _tc_ok, tensor_core_status
    monkeypatch.setattr(E, 'USE_DENSE_TC', True)
    g0 = torch.Generator().manual_seed(n + n_out)
    x = torch.randn((n, n_in), generator=g0).to(dev).requires_grad_(True)
    w = (torch.randn((n_out, n_in), generator=g0) * 0.3).to(dev).requires_grad_(True)
    b = torch.randn(n_out, generator=g0).to(dev).requires_grad_(True)
    go = torch.randn((n, n_out), generator=g0).to(dev)
    assert _tc_ok(n_in, n_out, x)
    y = _DenseFn.apply(x, w, b, relu)
    y.backward(go)
    x64, w64, b64 = (t.detach().double().requires_grad_(True) for t in (x, w, b))
    z = x64 @ w64.t() + b64
    ref = torch.relu(z) if relu else z
    # the kernel decides relu' from ITS OWN output; compare gradients with the same mask (TF32 can flip near-zero z)
    mask = (y.detach() > 0).double() if relu else torch.ones_like(z)
    scale = float(ref.abs().max())
    assert float((y.double() - ref).abs().max()) < 4e-3 * max(scale, 1.0)
    gz = go.double() * mask
    assert float((x.grad.double() - gz @ w64).abs().max()) < 4e-3 * max(float((gz @ w64).abs().max()), 1.0)
    assert float((w.grad.double() - gz.t() @ x64).abs().max()) < 4e-3 * max(float((gz.t() @ x64).abs().max()), 1.0)
    assert float((b.grad.double() - gz.sum(0)).abs().max()) < 4e-3 * max(float(gz.sum(0).abs().max()), 1.0)
    assert tensor_core_status(dev) == 0


@pytest.mark.parametrize('n,n_in,n_out', [(5000, 11, 60), (70001, 60, 60), (333, 60, 5), (31, 60, 11), (1, 63, 64)])
def test_dense_weight_gradient_kernel(qb, dev, n, n_in, n_out):
    """qbold_dense_wgrad (TF32 mma, fp32 accumulate) vs float64 matmul; the encoder routes its Dense layers through it."""
    from qbold_vi_b200.encoder import _DenseFn
    g0 = torch.Generator().manual_seed(n + n_in)
    x = torch.randn((n, n_in), generator=g0).to(dev).requires_grad_(True)
    w = (torch.randn((n_out, n_in), generator=g0) * 0.2).to(dev).requires_grad_(True)
    b = torch.randn(n_out, generator=g0).to(dev).requires_grad_(True)
    go = torch.randn((n, n_out), generator=g0).to(dev)
    y = _DenseFn.apply(x, w, b)
    y.backward(go)
    dw_ref = go.double().t() @ x.detach().double()
    db_ref = go.double().sum(0)
    scale = float(dw_ref.abs().max())
    # TF32 operands: 2^-11 relative per product, random signs -> ~1e-3 of the scale at most
    assert float((w.grad.double() - dw_ref).abs().max()) < 2e-3 * max(scale, 1.0)
    assert float((b.grad.double() - db_ref).abs().max()) < 2e-3 * max(float(db_ref.abs().max()), 1.0)
    assert float((x.grad.double() - go.double() @ w.detach().double()).abs().max()) < 5e-3 * float(go.abs().max())
    y2 = _DenseFn.apply(x, w, b)                               # deterministic: bit-identical on a second run
    w.grad = None
    y2.backward(go)
    dw1 = w.grad.clone()
    w.grad = None
    _DenseFn.apply(x, w, b).backward(go)
    assert torch.equal(dw1, w.grad)


@pytest.mark.parametrize('zc', [60, 1])
def test_gate_mix_kernels(qb, dev, zc):
    """Gated residual mix of the encoder blocks (model.py:160-172) vs the torch expression and its autograd."""
    from qbold_vi_b200.encoder import gate_mix
    g0 = torch.Generator().manual_seed(zc)
    shp = (2, 7, 5, 3, 60)
    skip, r = (torch.randn(shp, generator=g0).to(dev).requires_grad_(True) for _ in range(2))
    z = torch.randn(shp[:-1] + (zc,), generator=g0).to(dev).requires_grad_(True)
    go = torch.randn(shp, generator=g0).to(dev)
    out = gate_mix(skip, r, z, -3.0)
    out.backward(go)
    got = [out.detach().clone(), skip.grad.clone(), r.grad.clone(), z.grad.clone()]
    for t in (skip, r, z):
        t.grad = None
    g = torch.sigmoid(z.double() - 3.0)
    ref = skip.double() * (1.0 - g) + r.double() * g
    ref.backward(go.double())
    for a, b in zip(got, [ref.detach(), skip.grad, r.grad, z.grad]):
        assert float((a.double() - b.double()).abs().max()) < 1e-5 * max(float(b.abs().max()), 1.0)


def test_concurrent_streams_share_no_work_counter(qb, dev, cfg_noise_off):
    """Every launch takes its own device work counter (a ring, re-armed on the launch's stream): kernels overlapping
    on different streams, and > ring-size launches on one stream, must each cover all of their voxels exactly once."""
    layer = qb.SignalGenerationLayer(cfg_noise_off, True, True)
    xs = [_t(_rand_voxels(200_003 + 17 * i, 40 + i), dev) for i in range(4)]
    refs = [layer(x) for x in xs]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(device=dev) for _ in range(4)]
    outs = [[] for _ in range(4)]
    for rep in range(6):
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                outs[i].append(layer(xs[i]))
    torch.cuda.synchronize()
    for i in range(4):
        for o_ in outs[i]:
            assert torch.equal(o_, refs[i])
    small = _t(_rand_voxels(33, 3), dev)
    ref_small = layer(small)
    for _ in range(1100):                                   # wraps the 1024-entry ring on one stream
        got = layer(small)
    assert torch.equal(got, ref_small)


def test_cuda_graph_capture_and_replay(qb, dev, cfg_noise_off):
    """The launches (work-counter re-arm + kernel) are capturable: a CUDA graph of forward + VJP replays bit-identically
    on new inputs written into the captured buffers."""
    layer = qb.SignalGenerationLayer(cfg_noise_off, True, True)
    n = 50_001
    x = _t(_rand_voxels(n, 1), dev)
    g = torch.randn((n, 11), device=dev)
    sig, grad = torch.empty((n, 11), device=dev), torch.empty((n, 2), device=dev)
    import ctypes as C
    lib = qb._lib.lib()
    layer(x)                                                                 # warm-up outside capture
    s = torch.cuda.Stream(device=dev)
    s.wait_stream(torch.cuda.current_stream(dev))
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        qb._lib.check(lib.qbold_forward_backward(C.byref(layer.params), x.data_ptr(), g.data_ptr(), n, sig.data_ptr(),
                                                 grad.data_ptr(), qb._lib.stream_ptr(dev)))
        torch.cuda.synchronize()
        with torch.cuda.graph(graph, stream=s):
            qb._lib.check(lib.qbold_forward_backward(C.byref(layer.params), x.data_ptr(), g.data_ptr(), n, sig.data_ptr(),
                                                     grad.data_ptr(), qb._lib.stream_ptr(dev)))
    for seed in (2, 3):
        x.copy_(_t(_rand_voxels(n, seed), dev))
        g.copy_(torch.randn((n, 11), device=dev))
        graph.replay()
        torch.cuda.synchronize()
        s_ref, g_ref = layer.forward_backward(x, g)
        assert torch.equal(sig, s_ref) and torch.equal(grad, g_ref)


def test_c_abi_from_plain_c(qb, dev, tmp_path):
    """examples/c_abi_demo.c: the library driven from C (no Python, no torch) reproduces the Appendix B known answers."""
    import os
    import shutil
    import subprocess
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        pytest.skip('nvcc not available on this box')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, 'qbold_vi_b200')
    exe = str(tmp_path / 'c_abi_demo')
    subprocess.run([nvcc, '-Wno-deprecated-gpu-targets', '-x', 'c', os.path.join(root, 'examples', 'c_abi_demo.c'), '-I',
                    os.path.join(root, 'include'), '-L', libdir, '-lqbold', '-Xlinker', '-rpath=' + libdir, '-o', exe],
                   check=True, capture_output=True, timeout=300)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert 'max relative deviation' in out.stdout


def test_dlpack_hand_off_without_torch_tensors(qb, dev, cfg_noise_off):
    """north_star: "ctypes, with DLPack for tensor hand-off".  A foreign producer (here: a minimal object exposing only
    __dlpack__ / __dlpack_device__, and a raw capsule) reaches qbold_forward / qbold_forward_backward as raw pointers."""
    from torch.utils import dlpack as tdl

    class Foreign:                                                   # what a tf / cupy / jax array looks like to us
        def __init__(self, t):
            self._t = t

        def __dlpack__(self, stream=None):
            return self._t.__dlpack__(stream=stream)

        def __dlpack_device__(self):
            return self._t.__dlpack_device__()

    layer = qb.SignalGenerationLayer(cfg_noise_off, True, True)
    x = _t(_rand_voxels(3001, 23), dev)
    want_s, want_g = layer.forward_backward(x, None)
    out = torch.empty((3001, 11), device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    qb.forward_dlpack(layer, Foreign(x), tdl.to_dlpack(out), stream=st)
    assert torch.equal(out, want_s)
    sig, grad = torch.zeros((3001, 11), device=dev), torch.zeros((3001, 2), device=dev)
    qb.forward_backward_dlpack(layer, tdl.to_dlpack(x), None, Foreign(sig), Foreign(grad), stream=st)
    assert torch.equal(sig, want_s) and torch.equal(grad, want_g)
    assert torch.equal(layer(Foreign(x)), want_s)                    # the layer call itself takes producers too
    with pytest.raises(qb.dlpack.DLPackError):
        qb.forward_dlpack(layer, Foreign(x), Foreign(torch.empty((7, 11), device=dev)))
    with pytest.raises(qb.dlpack.DLPackError, match='CUDA device memory only'):
        qb.forward_dlpack(layer, x.cpu(), Foreign(out))


def test_fused_encoder_block_path_matches_the_layer_by_layer_path(qb, dev, monkeypatch):
    """The training path of the encoder (z-outer layout, one fused autograd node per gated residual block, bias of the
    second convolution folded into the gate Dense / mix kernel, fused ReLU' + bias-gradient kernels) against the plain
    layer-by-layer module: outputs and every parameter gradient, in strict float32 (TF32 off) so the bar can be tight."""
    import qbold_vi_b200.encoder as E
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        torch.manual_seed(3)
        enc = E.Encoder(no_units=60, no_intermediate_layers=2, gate_offset=-1.0, resid_init_std=0.1).to(dev)
        data = torch.rand(2, 10, 12, 6, 11, device=dev) * 50.0 + 20.0
        w = [torch.randn(2, 10, 12, 6, c, device=dev) for c in (5, 5, 11)]

        def run(fast):
            monkeypatch.setattr(E, '_FAST_BLOCK', fast)
            for p in enc.parameters():
                p.grad = None
            outs = enc(data)
            sum((o * wi).sum() for o, wi in zip(outs, w)).backward()
            return [o.detach().clone() for o in outs], [p.grad.detach().clone() for p in enc.parameters()]

        o_ref, g_ref = run(False)
        o_fast, g_fast = run(True)
        for a, b in zip(o_fast, o_ref):
            assert a.shape == b.shape and float((a - b).abs().max()) <= 2e-5 * float(b.abs().max())
        names = [n for n, _ in enc.named_parameters()]
        for n, a, b in zip(names, g_fast, g_ref):
            assert a.shape == b.shape, n
            assert float((a - b).abs().max()) <= 1e-4 * float(b.abs().max()) + 1e-6, n
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def test_fused_encoder_path_matches_the_reference_source_network(qb, dev):
    """The GPU training path (fused blocks, z-outer layout) against the reference-source encoder fixture, strict float32."""
    from conftest import load_reference_encoder_weights, reference_encoder_grad
    from qbold_vi_b200.encoder import Encoder
    fix = golden('ref_shim_encoder.npz')
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        enc = Encoder(no_units=60, no_intermediate_layers=2, activation='relu', initial_im_sigma=0.05,
                      multi_image_normalisation=False, channelwise_gating=True, gate_offset=float(fix['gate_offset']),
                      resid_init_std=0.1, no_ip_images=11, se_idx=2, use_mvg=True).to(dev)
        params = load_reference_encoder_weights(enc, fix)
        outs = enc(_t(fix['data'], dev))
        for o, key in zip(outs, ('out_voxelwise', 'out_spatial', 'out_sigma')):
            assert rel_max(o.detach().cpu().numpy(), fix[key]) < 5e-5, key
        sum((o * _t(fix['w_out%d' % i], dev)).sum() for i, o in enumerate(outs)).backward()
        for i, (w, b, kind) in enumerate(params):
            gw, gb = reference_encoder_grad(fix, i, kind)
            assert rel_max(w.grad.cpu().numpy(), gw.numpy()) < 2e-4, ('kernel', i)
            assert rel_max(b.grad.cpu().numpy(), gb.numpy()) < 2e-4, ('bias', i)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


@pytest.mark.parametrize('n,c', [(1, 4), (257, 60), (100003, 60), (4096, 64)])
def test_encoder_block_kernels_match_torch(qb, dev, n, c):
    """csrc/encoder_block.cu against plain float32 torch: gated mix with the folded convolution bias (+ ReLU copy), its
    backward with the skip branch's ReLU' applied in place, and ReLU' + column sums in one pass; ragged sizes."""
    import ctypes as C
    from qbold_vi_b200._lib import check, dptr, lib, stream_ptr
    g = torch.Generator(device=dev).manual_seed(n + c)
    rnd = lambda *s: torch.randn(*s, device=dev, generator=g)                       # noqa: E731
    skip, r0, z, go = torch.relu(rnd(n, c)), rnd(n, c), rnd(n, c), rnd(n, c)
    rb = rnd(c)
    off = -0.7
    out, out_relu = torch.empty_like(r0), torch.empty_like(r0)
    st = stream_ptr(dev)
    check(lib().qbold_block_mix_forward(dptr(skip), dptr(r0), dptr(rb), dptr(z), off, n, c, dptr(out), dptr(out_relu), st))
    gate = torch.sigmoid(z + off)
    want = skip * (1 - gate) + (r0 + rb) * gate
    assert torch.allclose(out, want, rtol=1e-6, atol=1e-6) and torch.allclose(out_relu, torch.relu(want), rtol=1e-6, atol=1e-6)
    ds, dr, dz = torch.empty_like(r0), torch.empty_like(r0), torch.empty_like(r0)
    check(lib().qbold_block_mix_backward(dptr(go), dptr(skip), dptr(r0), dptr(rb), dptr(z), off, n, c, 1, dptr(ds), dptr(dr),
                                         dptr(dz), st))
    assert torch.allclose(ds, go * (1 - gate) * (skip > 0), rtol=1e-6, atol=1e-6)
    assert torch.allclose(dr, go * gate, rtol=1e-6, atol=1e-6)
    assert torch.allclose(dz, go * ((r0 + rb) - skip) * gate * (1 - gate), rtol=1e-5, atol=1e-6)
    # ReLU' (+ addend) with the bias gradient from the same pass
    y, add = rnd(n, c), rnd(n, c)
    ws = torch.empty(int(lib().qbold_colsum_workspace_floats()), device=dev)
    o, cs = torch.empty_like(y), torch.empty(c, device=dev)
    check(lib().qbold_relu_bwd_colsum(dptr(go), dptr(y), dptr(add), n, c, dptr(o), dptr(cs), 0, dptr(ws), st))
    want = go * (y > 0) + add
    assert torch.equal(o, want)
    ref = want.double().sum(0)
    assert float((cs.double() - ref).abs().max()) <= 1e-5 * float(want.abs().double().sum(0).max()) + 1e-6
    cs2 = cs.clone()
    check(lib().qbold_relu_bwd_colsum(dptr(go), None, None, n, c, None, dptr(cs2), 1, dptr(ws), st))     # accumulate
    ref2 = ref + go.double().sum(0)
    assert float((cs2.double() - ref2).abs().max()) <= 1e-5 * float(go.abs().double().sum(0).max() + ref.abs().max()) + 1e-6
    with pytest.raises(qb.QboldError):
        check(lib().qbold_block_mix_forward(dptr(skip), dptr(r0), None, dptr(z), off, n, 6, dptr(out), None, st))   # c % 4


def test_new_entry_points_edge_cases(qb, dev, cfg_noise_off):
    """Empty and degenerate inputs of the round-2 entry points (the reference's behaviour for empty tensors is 'empty in,
    empty out'; bad shapes raise)."""
    from qbold_vi_b200._lib import check, dptr, lib, stream_ptr
    layer = qb.SignalGenerationLayer(cfg_noise_off, True, True, misaligned_prob=0.5)
    empty = torch.empty((0, 2), device=dev)
    assert tuple(layer.misalign(empty, torch.empty((0, 11), device=dev)).shape) == (0, 11)
    x = _t(_rand_voxels(64, 9), dev)
    s = layer._forward_raw(x)
    layer._misaligned_prob = 0.0
    assert torch.equal(layer.misalign(x, s), s)                                      # prob 0: untouched
    layer._misaligned_prob = 0.5
    with pytest.raises(qb.QboldError):                                               # all three draw arrays or none
        layer.misalign(x, s, sel_u01=torch.rand(64, device=dev))
    hl = qb.SignalGenerationLayer(cfg_noise_off, True, True, variable_hct=True)
    sig, g = hl.forward_backward(torch.empty((0, 3), device=dev), None)
    assert tuple(sig.shape) == (0, 11) and tuple(g.shape) == (0, 3)
    one = torch.tensor([[0.4, 0.12, 0.34]], device=dev)
    s1, g1 = hl.forward_backward(one, None)
    s2, g2 = qb.SignalGenerationLayer(cfg_noise_off, True, True).forward_backward(one[:, :2].contiguous(), None)
    assert torch.allclose(s1, s2, rtol=2e-6, atol=0) and torch.allclose(g1[:, :2], g2, rtol=1e-5, atol=0)
    tr = _trainer(qb, cfg_noise_off, use_mvg=False, use_population_prior=True, mog_components=2)
    q12 = torch.randn(5, 12, device=dev)
    dead = torch.zeros(5, 5, device=dev)                                             # mask channel (index 4) all zero
    kl = tr.kl_loss(dead, q12, return_mean=False)
    assert tuple(kl.shape) == (5, 1) and float(kl.abs().max()) == 0.0
    with pytest.raises(ValueError):
        tr.kl_loss(dead, torch.randn(5, 8, device=dev))                              # wrong channel count for M = 2


@pytest.mark.parametrize('bz,nx,ny,cg,cx', [(3, 5, 7, 60, 60), (2, 64, 64, 60, 60), (1, 1, 1, 4, 8), (4, 9, 33, 12, 8),
                                           (130, 16, 16, 64, 64)])
def test_conv_weight_gradient_tensor_core_kernel(qb, dev, bz, nx, ny, cg, cx):
    """qbold_conv_wgrad (tcgen05 kind::tf32, voxel-contraction GEMM with nine shifted operands) against the float64
    weight gradient of a 3x3 'same' convolution; TF32 operand rounding sets the bar (1e-3 of the largest entry)."""
    from qbold_vi_b200._lib import check, dptr, lib, stream_ptr
    gen = torch.Generator(device=dev).manual_seed(bz * 1000 + nx)
    x = torch.randn(bz, nx, ny, cx, device=dev, generator=gen)
    g = torch.randn(bz, nx, ny, cg, device=dev, generator=gen)
    w = torch.zeros(cg, cx, 3, 3, device=dev, dtype=torch.float64, requires_grad=True)
    y = torch.nn.functional.conv2d(x.double().permute(0, 3, 1, 2), w, None, padding=1)
    (y * g.double().permute(0, 3, 1, 2)).sum().backward()
    dw = torch.full((cg, cx, 3, 3), float('nan'), device=dev)
    ws = torch.empty(int(lib().qbold_conv_wgrad_workspace_floats()), device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    check(lib().qbold_conv_wgrad(dptr(g.reshape(-1, cg)), cg, dptr(x.reshape(-1, cx)), cx, bz, nx, ny, dptr(dw), 0, dptr(ws),
                                 dptr(status, torch.int32), stream_ptr(dev)))
    assert int(status.item()) == 0
    ref = w.grad
    scale = float(ref.abs().max()) + float((bz * nx * ny) ** 0.5) * 1e-3
    assert float((dw.double() - ref).abs().max()) <= 1.5e-3 * scale
    dw2 = dw.clone()
    check(lib().qbold_conv_wgrad(dptr(g.reshape(-1, cg)), cg, dptr(x.reshape(-1, cx)), cx, bz, nx, ny, dptr(dw2), 1, dptr(ws),
                                 dptr(status, torch.int32), stream_ptr(dev)))
    assert torch.allclose(dw2, 2 * dw, rtol=1e-6, atol=1e-6)                          # accumulate, deterministic
    with pytest.raises(qb.QboldError):
        check(lib().qbold_conv_wgrad(dptr(g.reshape(-1, cg)), 6, dptr(x.reshape(-1, cx)), cx, bz, nx, ny, dptr(dw), 0, dptr(ws),
                                     None, stream_ptr(dev)))


@pytest.mark.parametrize('n,n_in,n_out,transpose,relu,add', [(1000, 60, 60, 0, 1, 0), (4097, 60, 60, 1, 0, 1), (129, 32, 64, 0, 0, 1),
                                                            (1, 64, 60, 1, 1, 0), (70000, 60, 60, 0, 0, 0), (300, 8, 12, 0, 1, 0),
                                                            (555, 12, 44, 1, 0, 1)])
def test_dense_tma_pipeline_kernel(qb, dev, n, n_in, n_out, transpose, relu, add):
    """qbold_dense_tma (TMA -> tcgen05 kind::tf32 -> TMA store; forward with the stored weight K-major, input gradient
    with the same rows read MN-major; bias, ReLU and the in-place beta = 1 addend fused) against float64; TF32 operand
    rounding sets the bar."""
    from qbold_vi_b200._lib import check, dptr, lib, stream_ptr
    gen = torch.Generator(device=dev).manual_seed(n + n_in)
    x = torch.randn(n, n_in, device=dev, generator=gen)
    w = torch.randn((n_in, n_out) if transpose else (n_out, n_in), device=dev, generator=gen) * 0.3
    bias = None if transpose else torch.randn(n_out, device=dev, generator=gen)
    y = torch.randn(n, n_out, device=dev, generator=gen) if add else torch.full((n, n_out), float('nan'), device=dev)
    ref = x.double() @ (w.double() if transpose else w.double().t())
    if bias is not None:
        ref = ref + bias.double()
    if relu:
        ref = ref.clamp_min(0)
    if add:
        ref = ref + y.double()
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    check(lib().qbold_dense_tma(dptr(x), dptr(w), dptr(bias, allow_none=True), dptr(y) if add else None, n_in, n_out, transpose,
                                relu, n, dptr(y), dptr(status, torch.int32), stream_ptr(dev)))
    assert int(status.item()) == 0
    scale = float(ref.abs().max()) + 1e-6
    assert float((y.double() - ref).abs().max()) <= 2e-3 * scale
    with pytest.raises(qb.QboldError):
        check(lib().qbold_dense_tma(dptr(x), dptr(w), None, None, 6, n_out, 0, 0, n, dptr(y), None, stream_ptr(dev)))


@pytest.mark.parametrize('n,n_in,n_out', [(1000, 60, 60), (70001, 60, 60), (1, 64, 64), (129, 12, 44), (5000, 32, 8)])
def test_dense_weight_gradient_tma_kernel(qb, dev, n, n_in, n_out):
    """qbold_dense_wgrad_tma (g and x read MN-major by TMA, contraction over the rows on tcgen05, bias gradient from an
    MMA against a tile of ones) against float64, plus accumulation and determinism."""
    from qbold_vi_b200._lib import check, dptr, lib, stream_ptr
    gen = torch.Generator(device=dev).manual_seed(n + n_out)
    x = torch.randn(n, n_in, device=dev, generator=gen)
    g = torch.randn(n, n_out, device=dev, generator=gen)
    dw = torch.full((n_out, n_in), float('nan'), device=dev)
    db = torch.full((n_out,), float('nan'), device=dev)
    ws = torch.empty(int(lib().qbold_dense_wgrad_tma_workspace_floats()), device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    args = (dptr(g), n_out, dptr(x), n_in, n)
    check(lib().qbold_dense_wgrad_tma(*args, dptr(dw), dptr(db), 0, dptr(ws), dptr(status, torch.int32), stream_ptr(dev)))
    assert int(status.item()) == 0
    ref_w, ref_b = g.double().t() @ x.double(), g.double().sum(0)
    scale = float(ref_w.abs().max()) + float(n ** 0.5) * 1e-3
    assert float((dw.double() - ref_w).abs().max()) <= 1.5e-3 * scale
    assert float((db.double() - ref_b).abs().max()) <= 1.5e-3 * (float(ref_b.abs().max()) + float(n ** 0.5) * 1e-3)
    dw2, db2 = dw.clone(), db.clone()
    check(lib().qbold_dense_wgrad_tma(*args, dptr(dw2), dptr(db2), 1, dptr(ws), dptr(status, torch.int32), stream_ptr(dev)))
    assert torch.allclose(dw2, 2 * dw, rtol=1e-6, atol=1e-6) and torch.allclose(db2, 2 * db, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize('shape,multi', [((2, 5, 7, 9, 11), False), ((1, 3, 40, 33, 11), True), ((1, 2, 64, 64, 24), False)])
def test_normalise_zouter_kernel(qb, dev, shape, multi):
    """qbold_normalise_zouter = Encoder.normalise_data (model.py:97-113) + the permute to z-outer rows, zero padded to a
    multiple of 4 images."""
    from qbold_vi_b200._lib import check, dptr, lib, stream_ptr
    from qbold_vi_b200.encoder import Encoder
    b, nx, ny, nz, t = shape
    gen = torch.Generator(device=dev).manual_seed(sum(shape))
    data = torch.rand(shape, device=dev, generator=gen) * 3.0 - 0.2              # some values below the 1e-2 clip
    enc = Encoder(no_ip_images=t, se_idx=2, multi_image_normalisation=multi)
    ref = enc.normalise_data(data).permute(0, 3, 1, 2, 4).reshape(-1, t)
    tp = (t + 3) & ~3
    out = torch.full((ref.shape[0], tp), float('nan'), device=dev)
    check(lib().qbold_normalise_zouter(dptr(data), b, nx, ny, nz, t, 2, int(multi), dptr(out), stream_ptr(dev)))
    assert torch.allclose(out[:, :t], ref, rtol=1e-6, atol=2e-6)
    assert float(out[:, t:].abs().sum()) == 0.0


@pytest.mark.parametrize('n,n_in,n_out,masked', [(1000, 60, 5, True), (33, 60, 16, False), (70001, 60, 5, True), (1, 12, 3, True)])
def test_skinny_input_gradient_with_relu_mask(qb, dev, n, n_in, n_out, masked):
    """qbold_dense_small_dgrad_masked: dx = [mask > 0] * (g W) in float32 (warp-cooperative stores)."""
    from qbold_vi_b200._lib import check, dptr, lib, stream_ptr
    gen = torch.Generator(device=dev).manual_seed(n + n_out)
    g = torch.randn(n, n_out, device=dev, generator=gen)
    w = torch.randn(n_out, n_in, device=dev, generator=gen)
    act = torch.relu(torch.randn(n, n_in, device=dev, generator=gen))
    dx = torch.full((n, n_in), float('nan'), device=dev)
    check(lib().qbold_dense_small_dgrad_masked(dptr(g), dptr(w), dptr(act) if masked else None, n_in, n_out, n, dptr(dx),
                                               stream_ptr(dev)))
    ref = g.double() @ w.double()
    if masked:
        ref = ref * (act > 0)
    assert float((dx.double() - ref).abs().max()) <= 1e-5 * (float(ref.abs().max()) + 1.0)


def test_fused_encoder_production_path_in_tf32(qb, dev, monkeypatch):
    """The encoder as training runs it (TF32 on: normalise + z-outer kernel, Dense layers and weight gradients on the
    TMA / tcgen05 kernels, masked head gradient, cuDNN TF32 convolutions) against the layer-by-layer module in strict
    float32: outputs within 1e-2, every parameter gradient within twice the deviation the layer-by-layer module shows when
    the libraries run it in TF32."""
    import qbold_vi_b200.encoder as E
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    try:
        torch.manual_seed(5)
        enc = E.Encoder(no_units=60, no_intermediate_layers=2, gate_offset=-1.0, resid_init_std=0.1).to(dev)
        data = torch.rand(2, 18, 20, 7, 11, device=dev) * 50.0 + 20.0
        w = [torch.randn(2, 18, 20, 7, c, device=dev) for c in (5, 5, 11)]

        def run(fast, tf32):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            monkeypatch.setattr(E, '_FAST_BLOCK', fast)
            for p in enc.parameters():
                p.grad = None
            outs = enc(data)
            sum((o * wi).sum() for o, wi in zip(outs, w)).backward()
            return [o.detach().clone() for o in outs], [p.grad.detach().clone() for p in enc.parameters()]

        o_ref, g_ref = run(False, False)
        o_lib, g_lib = run(False, True)                                 # the library's own TF32 error on the same network
        launches = qb.launch_count()
        o_fast, g_fast = run(True, True)
        assert qb.launch_count() - launches >= 30                      # the hand-written kernels ran, not a library fallback
        assert E.tensor_core_status(dev) == 0
        o_fast[2], o_ref[2] = torch.log(o_fast[2]), torch.log(o_ref[2])      # sigma = exp(.): compare the exponent
        errs = [float((a - b).abs().max()) / float(b.abs().max()) for a, b in zip(o_fast, o_ref)]
        gerrs = {n: float((a - b).abs().max()) / (float(b.abs().max()) + 1e-6)
                 for (n, _), a, b in zip(enc.named_parameters(), g_fast, g_ref)}
        lerrs = {n: float((a - b).abs().max()) / (float(b.abs().max()) + 1e-6)
                 for (n, _), a, b in zip(enc.named_parameters(), g_lib, g_ref)}
        print('tf32 path: output errors', errs, 'worst gradient', max(gerrs.items(), key=lambda kv: kv[1]),
              'library TF32 worst gradient', max(lerrs.items(), key=lambda kv: kv[1]))
        for a, b in zip(o_fast, o_ref):
            assert a.shape == b.shape
        assert max(errs) <= 1e-2, errs
        # gradients: random cotangents make them sums of cancelling terms, so TF32 rounding shows up at the per-cent
        # level in ANY TF32 implementation; the bar is the library path's own deviation from float32 on the same network
        for n in gerrs:
            assert gerrs[n] <= 2.0 * lerrs[n] + 1e-2, (n, gerrs[n], lerrs[n])
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


@pytest.mark.parametrize('n', [1, 300001, (1 << 20) + (1 << 18) + 777])
def test_host_buffer_pipeline_matches_the_device_path(qb, dev, cfg_noise_off, n):
    """qbold_forward_backward_host (pinned host buffers, chunked H2D / kernel / D2H pipeline with super-chunked OEF/DBV
    and gradient copies) gives bit for bit what the device entry point gives, across chunk and super-chunk boundaries
    and on a second call that reuses the pipeline's buffers."""
    layer = qb.SignalGenerationLayer(cfg_noise_off, True, True)
    gen = torch.Generator().manual_seed(n)
    hx = (torch.rand(n, 2, generator=gen) * torch.tensor([0.8, 0.2]) + torch.tensor([0.04, 0.001])).pin_memory()
    hg = torch.randn(n, 11, generator=gen).pin_memory()
    s_ref, g_ref = layer.forward_backward(hx.to(dev), hg.to(dev))
    for _ in range(2):
        hs = torch.full((n, 11), float('nan')).pin_memory()
        hgr = torch.full((n, 2), float('nan')).pin_memory()
        layer.forward_backward_host(hx, hg, hs, hgr)
        assert torch.equal(hs, s_ref.cpu()) and torch.equal(hgr, g_ref.cpu())


@pytest.mark.xfail(strict=False, reason='split mode (collectives outside the capture) was written after the GPU budget of '
                   'round 2 was spent: not yet run on hardware')
def test_split_captured_training_step_matches_the_eager_step(qb, dev, cfg_noise_off, tmp_path):
    """cuda_graph='split': two graphs (encoder + losses + backward | weight decay + Adam) with the collectives enqueued
    eagerly around them -- the multi-GPU form of the captured step.  Last in the file on purpose."""
    _captured_step_against_eager(qb, dev, cfg_noise_off, tmp_path, 'split')
