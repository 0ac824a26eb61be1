"""CPU suite, part 1: pin the oracle.

(a) against tests/golden/ref_shim_*.npz -- outputs of the reference's own, unmodified
    signals.py / model.py executed over the TensorFlow API shim (oracle/make_golden.py);
(b) against the known-answer vectors of SURVEY.md Appendix B;
(c) structural identities of the model.
Tolerances: float32-emulated oracle vs fixtures 2e-6 (same arithmetic, different op library);
float64 oracle vs fixtures 1e-5 (the FP32-vs-FP64 gap measured in SURVEY.md A.6 is 2.9e-6).
"""
import numpy as np
import pytest

from conftest import golden, rel_elem, rel_max
from oracle import qbold_oracle as o


def test_tau_grid_and_constants(physics):
    g = golden('ref_shim_forward.npz')
    assert np.array_equal(g['taus'], physics.taus)
    assert physics.taus[2] == 0.0 and physics.n_tau == 11
    assert abs(o.dw_const(physics) - 301.743275) < 1e-5
    assert abs(float(o._exp(-physics.te * physics.r2t, np.float32)) - 0.42698773) < 1e-7
    assert abs(float(o.m_bld(physics, np.float64)) - 0.21985563) < 2e-7


@pytest.mark.parametrize('full', [1, 0])
@pytest.mark.parametrize('blood', [1, 0])
@pytest.mark.parametrize('dt,tol', [(np.float32, 2e-6), (np.float64, 1e-5)])
def test_forward_and_tape_gradient_vs_reference_source(physics, full, blood, dt, tol):
    g = golden('ref_shim_forward.npz')
    key = 'f%d_b%d' % (full, blood)
    x = g['oef_dbv']
    S = o.forward(physics, x, bool(full), bool(blood), dt)
    assert rel_elem(S, g['signal_' + key]) < tol
    _, g1 = o.forward_backward(physics, x, np.ones((x.shape[0], 11)), bool(full), bool(blood), dt)
    _, g2 = o.forward_backward(physics, x, g['g_rand'], bool(full), bool(blood), dt)
    assert rel_max(g1, g['grad_ones_' + key]) < tol
    assert rel_max(g2, g['grad_rand_' + key]) < tol


def test_appendix_b_known_answers(physics):
    k = golden('kat_appendix_b.npz')
    s64 = o.forward(physics, k['oef_dbv'], dtype=np.float64)
    s32 = o.forward(physics, k['oef_dbv'], dtype=np.float32)
    assert np.max(np.abs(s64[0] - k['signal_fp64_0'])) < 5e-9
    assert np.max(np.abs(s32[1] - k['signal_fp32_1'])) < 5e-8
    assert np.max(np.abs(s32 - s64)) < 4e-7
    _, g = o.forward_backward(physics, k['oef_dbv'][:1], np.ones((1, 11)), dtype=np.float64)
    assert np.max(np.abs(g[0] - k['grad_sum_0'])) < 5e-8


@pytest.mark.parametrize('tag', ['optimal', 'multinorm'])
@pytest.mark.parametrize('dt,tol', [(np.float32, 3e-6), (np.float64, 2e-5)])
def test_elbo_and_gradients_vs_reference_source(physics, tag, dt, tol):
    e = golden('ref_shim_elbo_%s.npz' % tag)
    mn = bool(e['multi_image_normalisation'])
    smp, _ = o.reparam_sample(e['q'], e['eps'], True, dt)
    assert rel_elem(smp, e['sampled']) < tol
    r = o.elbo_and_grads(physics, e['q'], e['sigma'], e['data'], e['mask'], e['prior'], e['eps'], e['eps_kl'], dt,
                         multi_image_normalisation=mn)
    r0 = o.elbo_and_grads(physics, e['q'], e['sigma'], e['data'], e['mask'], e['prior'], e['eps'], None, dt,
                          multi_image_normalisation=mn)
    assert rel_elem(r['pred'], e['pred']) < tol
    assert rel_elem(r['nll'], e['nll']) < tol
    assert rel_elem(r['kl'], e['kl']) < tol
    assert rel_max(r['nll_map'], e['nll_map']) < tol
    assert rel_max(r0['grad_q'], e['grad_q_nll']) < tol
    assert rel_max(r['grad_sigma'], e['grad_sigma']) < tol
    assert rel_max(r['grad_q'] - r0['grad_q'], e['grad_q_kl']) < 10 * tol
    kl2 = np.where(e['mask'] > 0, o.mc_kl(e['prior'], e['q'], e['eps_kl2'], dt), 0)
    assert rel_max(kl2, e['kl_map2']) < tol


def test_student_t_likelihood_vs_reference_source(physics):
    e = golden('ref_shim_elbo_studentt.npz')
    yt = np.concatenate([e['data'], e['mask'][:, None]], -1)
    for dt in (np.float32, np.float64):
        nll = o.fine_tune_nll(yt, e['pred'], e['sigma'], int(e['se_idx']), dt, False, False, float(e['student_t_df']))
        assert rel_elem(nll, e['nll']) < 2e-6


def test_posterior_stats_vs_reference_source(physics):
    m = golden('ref_shim_means.npz')
    for dt in (np.float32, np.float64):
        mu, var = o.posterior_stats(physics, m['q'], m['eps'], dt)
        assert rel_elem(mu, m['means']) < 2e-6
        assert rel_max(var, m['stds']) < 2e-6


@pytest.mark.parametrize('tag', ['u10', 'u0'])
def test_synthetic_dataset_vs_reference_source(tag):
    d = golden('ref_shim_dataset_%s.npz' % tag)
    cfg = o.default_config()                                   # simulate_noise = True, as in the INI
    ph = o.parse_params(cfg)
    S, up = int(d['sample_size']), float(d['uniform_prop'])
    f32 = np.float32
    # marginals from the recorded draws, exactly as signals.py:255-268 forms them
    oefs = np.concatenate([d['oef_u01'] * f32(0.8 - 0.05) + f32(0.05),
                           np.clip(d['oef_n01'] * f32(0.2) + f32(0.4), f32(0.05), f32(0.8))]).astype(f32)
    import scipy.stats as st
    a, b = (0.003 - 0.025) / 0.02, (0.195 - 0.025) / 0.02
    dbvs = np.concatenate([d['dbv_u01'] * f32(0.195 - 0.003) + f32(0.003),
                           st.truncnorm.ppf(d['dbv_tn_u01'], a, b, loc=0.025, scale=0.02).astype(f32)]).astype(f32)
    assert oefs.shape[0] == S and dbvs.shape[0] == S
    snr_u = (d['snr_u01'] * f32(120 - 50) + f32(50)).astype(f32)
    x, y = o.synthetic_dataset_from_draws(ph, oefs, dbvs, d['perm'], snr_u, d['noise_eps'])
    assert x.shape == d['train_x'].shape == (520, 11) and y.shape == d['train_y'].shape == (529, 3)
    assert rel_elem(y, d['train_y']) < 1e-6
    assert rel_elem(x, d['train_x']) < 5e-6


def _marginals_from_labels(d):
    """Un-shuffle the labels of a dataset fixture to recover the 23 OEF / 23 DBV marginals the reference drew."""
    grid = np.empty((529, 2), np.float32)
    grid[d['perm']] = d['train_y'][:, :2]
    g = grid.reshape(23, 23, 2)
    return g[:, 0, 0].copy(), g[0, :, 1].copy()


def test_synthetic_dataset_with_misalignment_vs_reference_source():
    """create_synthetic_dataset(..., misaligned_prob=0.3): the reference's own draws (uniform, randint, two normals per
    chunk, then the noise draws) replayed through the restatement of signals.py:80-96."""
    d = golden('ref_shim_dataset_misalign.npz')
    ph = o.parse_params(o.default_config())
    oefs, dbvs = _marginals_from_labels(d)
    f32 = np.float32
    snr_u = (d['snr_u01'] * f32(120 - 50) + f32(50)).astype(f32)
    mis = (float(d['prob']), d['mis_u01'], d['mis_index'], d['mis_eps'])
    x, y = o.synthetic_dataset_from_draws(ph, oefs, dbvs, d['perm'], snr_u, d['noise_eps'], misalign=mis)
    assert rel_elem(y, d['train_y']) < 1e-6
    assert rel_elem(x, d['train_x']) < 5e-6
    assert 100 < int((d['mis_u01'] < d['prob']).sum()) < 220 and d['mis_index'].min() >= 4 and d['mis_index'].max() <= 9
    x0, _ = o.synthetic_dataset_from_draws(ph, oefs, dbvs, d['perm'], snr_u, d['noise_eps'])
    assert rel_elem(x0, d['train_x']) > 1e-2                   # the augmentation is visible


def test_variable_hct_vs_reference_source(physics):
    g = golden('ref_shim_forward_hct.npz')
    x = g['oef_dbv_hct']
    for full in (1, 0):
        for blood in (1, 0):
            key = 'f%d_b%d' % (full, blood)
            for dt, tol in ((np.float32, 2e-6), (np.float64, 1e-5)):
                S = o.forward(physics, x, bool(full), bool(blood), dt, variable_hct=True)
                assert rel_elem(S, g['signal_' + key]) < tol
                _, g1 = o.forward_backward(physics, x, np.ones((x.shape[0], 11)), bool(full), bool(blood), dt,
                                           variable_hct=True)
                _, g2 = o.forward_backward(physics, x, g['g_rand'], bool(full), bool(blood), dt, variable_hct=True)
                for j in range(3):                              # OEF, DBV and Hct columns on their own scales
                    assert rel_max(g1[:, j], g['grad_ones_' + key][:, j]) < tol
                    assert rel_max(g2[:, j], g['grad_rand_' + key][:, j]) < 5 * tol
    sm = o.forward_misaligned(physics, x, float(g['mis_prob']), g['mis_u01'], g['mis_index'], g['mis_eps'],
                              dtype=np.float32, variable_hct=True)
    assert rel_elem(sm, g['signal_misaligned']) < 2e-6
    d = golden('ref_shim_dataset_hct.npz')
    ph = o.parse_params(o.default_config())
    oefs, dbvs = _marginals_from_labels(d)
    snr_u = (d['snr_u01'] * np.float32(70) + np.float32(50)).astype(np.float32)
    xs, ys = o.synthetic_dataset_from_draws(ph, oefs, dbvs, d['perm'], snr_u, d['noise_eps'], variable_hct=True)
    assert rel_elem(ys, d['train_y']) < 1e-6 and rel_elem(xs, d['train_x']) < 5e-6


# ------------------------------------------------------------------ structural identities
def test_tau_symmetry_and_tau0_column(physics):
    rng = np.random.default_rng(3)
    x = np.stack([rng.uniform(0.04, 0.84, 64), rng.uniform(0.001, 0.201, 64)], -1)
    st = o.calc_tissue(physics, x[:, 0], x[:, 1], dtype=np.float64)
    assert np.allclose(st[:, 0], st[:, 4], rtol=1e-6) and np.allclose(st[:, 1], st[:, 3], rtol=1e-6)
    assert np.allclose(st[:, 2], 0.42698773, atol=2e-8)         # tissue term at tau=0 == exp(-TE*R2t)
    d = 1e-6
    st2 = o.calc_tissue(physics, x[:, 0], x[:, 1] + d, dtype=np.float64)
    assert np.all(st2 <= st + 1e-12)                            # monotone decreasing in DBV


def test_node0_is_value_dead_but_gradient_live(physics):
    """FD of the forward disagrees with the TF gradient because node 0 is dead in the value."""
    x = np.array([[0.4, 0.12]])
    _, g = o.forward_backward(physics, x, np.ones((1, 11)), dtype=np.float64)
    h = 1e-6
    fd = (o.forward(physics, x + [[h, 0]], dtype=np.float64).sum() -
          o.forward(physics, x - [[h, 0]], dtype=np.float64).sum()) / (2 * h)
    assert abs(fd - (-3.0374)) < 2e-4 and abs(g[0, 0] - (-3.06411134)) < 1e-7
    assert abs(fd - g[0, 0]) / abs(g[0, 0]) > 5e-3


def test_logit_roundtrip_and_mc_kl_converges_to_closed_form():
    z = np.linspace(-9.5, 9.5, 101)[:, None] * np.ones((1, 2))
    zz = o.backwards_transform(o.forward_transform(z, np.float64), True, np.float64)
    assert np.max(np.abs(zz - z)) < 1e-6
    rng = np.random.default_rng(0)
    q = np.array([[-0.3, 0.2, -1.0, -0.1, 0.7]])
    p = np.array([[-0.5, 0.4, -1.3, 0.1, -0.6]])
    eps = rng.standard_normal((1, 400000, 2))
    kl_mc = o.mc_kl(p, q, eps, np.float64)[0]
    kl_cf = o.closed_form_kl(p, q)[0]
    assert abs(kl_mc - kl_cf) / kl_cf < 5e-3


def test_masked_voxels_contribute_nothing(physics):
    e = golden('ref_shim_elbo_optimal.npz')
    r = o.elbo_and_grads(physics, e['q'], e['sigma'], e['data'], e['mask'], e['prior'], e['eps'], e['eps_kl'][:, :4],
                         np.float64)
    dead = e['mask'] == 0
    assert dead.any()
    assert np.all(r['grad_q'][dead] == 0) and np.all(r['grad_sigma'][dead] == 0)
    assert np.all(r['nll_map'][dead] == 0) and np.all(r['kl_map'][dead] == 0)


# ---------------------------------------------------------------- losses either side of the path
def test_smoothness_oracle_vs_reference_source():
    a = golden('ref_shim_adjacent.npz')
    shp = tuple(int(v) for v in a['shape'])
    for tag, c in (('mvg', 5), ('diag', 4)):
        val, g = o.smoothness_loss(a['q5'][:, :c].reshape(shp + (c,)), a['mask'].reshape(shp))
        assert rel_elem(val, a['tv_' + tag]) < 1e-5
        assert rel_max(g.reshape(-1, c), a['tv_%s_grad' % tag]) < 1e-4


@pytest.mark.parametrize('tag,use_mvg,ig', [('mvg', True, (0.0, 0.0)), ('mvg_ig', True, (3.0, 0.15)),
                                            ('diag', False, (0.0, 0.0)), ('diag_ig', False, (3.0, 0.15))])
def test_synthetic_data_loss_oracle_vs_reference_source(tag, use_mvg, ig):
    a = golden('ref_shim_adjacent.npz')
    c = 5 if use_mvg else 4
    # float32 like the reference: row 0 sits on the clip, where 1 - 1e-6 is not representable (1 - x = 1.013e-6)
    rows, g = o.synthetic_data_nll(a['synth_%s_labels' % tag], a['q5'][:, :c], use_mvg, ig[0], ig[1], dt=np.float32)
    assert rel_elem(rows.astype(np.float64).mean(), a['synth_' + tag]) < 1e-5
    assert rel_max(g / rows.shape[0], a['synth_%s_grad' % tag]) < 1e-5
    rows64, g64 = o.synthetic_data_nll(a['synth_%s_labels' % tag], a['q5'][:, :c], use_mvg, ig[0], ig[1])
    assert rel_max(g64[1:], g[1:]) < 1e-4 and rel_max(rows64[1:], rows[1:]) < 1e-5


def test_diag_kl_oracle_vs_reference_source():
    a = golden('ref_shim_adjacent.npz')
    kl, gq, _ = o.diag_kl(a['prior5'][:, :4], a['q5'][:, :4])
    m = a['mask']
    assert rel_elem((kl * (m > 0)).sum() / m.sum(), a['kl_diag']) < 1e-5
    assert rel_max(gq * (m > 0)[:, None] / m.sum(), a['kl_diag_grad']) < 1e-4


def test_single_precision_bessel_restatement_against_scipy():
    """oracle/cephes_f32.py restates the Cephes j0f / j1f routines TensorFlow's float32 bessel_j0 / bessel_j1 dispatch to
    (signals.py:170 and its registered gradient).  Pinned here against scipy's double-precision J0 / J1: the arguments
    of the qBOLD path, 1.5 * tau * dw * u, stay below ~25; the asymptotic branch is checked far beyond that."""
    import scipy.special as sp
    from oracle import cephes_f32 as c
    x = np.linspace(0.0, 30.0, 300001).astype(np.float32)
    assert c.j0f(x).dtype == np.float32 and c.j1f(x).dtype == np.float32
    assert np.abs(c.j0f(x).astype(np.float64) - sp.j0(x.astype(np.float64))).max() < 2.5e-7
    assert np.abs(c.j1f(x).astype(np.float64) - sp.j1(x.astype(np.float64))).max() < 2.5e-7
    x = np.linspace(30.0, 400.0, 300001).astype(np.float32)
    assert np.abs(c.j0f(x).astype(np.float64) - sp.j0(x.astype(np.float64))).max() < 1e-6
    assert np.abs(c.j1f(x).astype(np.float64) - sp.j1(x.astype(np.float64))).max() < 1e-6
    # known answers: J0(0) = 1, J1(0) = 0, the first zeros of J0 and J1, and J0' = -J1 by central differences
    assert c.j0f(np.float32([0.0]))[0] == 1.0 and c.j1f(np.float32([0.0]))[0] == 0.0
    assert np.abs(c.j0f(np.float32([2.404825557695773, 5.520078110286311, 8.653727912911013]))).max() < 2e-7
    assert np.abs(c.j1f(np.float32([3.8317059702075125, 7.015586669815619]))).max() < 2e-7
    xm = np.linspace(0.5, 25.0, 2001)
    h = 1e-2
    d = (c.j0f((xm + h).astype(np.float32)).astype(np.float64) - c.j0f((xm - h).astype(np.float32)).astype(np.float64)) / \
        ((xm + h).astype(np.float32).astype(np.float64) - (xm - h).astype(np.float32).astype(np.float64))
    assert np.abs(d + c.j1f(xm.astype(np.float32))).max() < 5e-5


def test_cpu_baseline_port_computes_the_reference_path():
    """oracle/torch_port.py is what `bench.py --impl reference` and the `cpu_baseline` leg time.  It must be the same
    computation: signal and autodiff gradient against the fixture recorded from the reference's own signals.py, and
    against the float64 oracle on random voxels of the bench workload (float32 program: 2e-5 element-wise)."""
    import torch
    from oracle import torch_port as tp
    cfg = o.default_config()
    cfg['simulate_noise'] = 'False'
    ph = o.parse_params(cfg)
    d = golden('ref_shim_forward.npz')
    x = d['oef_dbv'].reshape(-1, 2).astype(np.float32)
    g = d['g_rand'].reshape(-1, 11).astype(np.float32)
    s, gr = tp.forward_backward(ph, torch.from_numpy(x), torch.from_numpy(g))
    want_s, want_g = d['signal_f1_b1'].reshape(-1, 11), d['grad_rand_f1_b1'].reshape(-1, 2)
    assert np.max(np.abs(s.numpy() - want_s) / np.abs(want_s)) < 2e-5
    assert np.abs(gr.numpy() - want_g).max() < 2e-5 * np.abs(want_g).max()
    rng = np.random.default_rng(0)
    n = 3000
    x = np.stack([rng.uniform(0.04, 0.84, n), rng.uniform(0.001, 0.201, n)], -1).astype(np.float32)
    g = rng.standard_normal((n, 11)).astype(np.float32)
    s, gr = tp.forward_backward(ph, torch.from_numpy(x), torch.from_numpy(g), chunk=1024)
    s64, g64 = o.forward_backward(ph, x, g, dtype=np.float64)
    assert np.max(np.abs(s.numpy() - s64) / np.abs(s64)) < 2e-5
    assert np.abs(gr.numpy() - g64).max() < 2e-5 * np.abs(g64).max()
