"""CPU suite, part 4: the forward kernels of the hot path executed on the HOST.

tests/host_emu compiles qbold_vi_b200/csrc/forward_kernels.cuh (with qbold_core.cuh, bessel.cuh, rng.cuh -- the source
libqbold.so is built from; the few PTX statements have an IEEE-meaning branch under QB_HOST_EMU, and the library's SASS
is byte-identical with and without those guards) with g++ and runs the kernels in a small SIMT emulator: one host
thread per CUDA thread, warp primitives at per-warp barriers, __shared__ as statics.  The parameter block comes from the
real qbold_params_init.  So the lane schedule, the phase flushes through shared memory, the shuffles, the work counter
and the two-voxel pairing of the headline kernel are compared with the oracle and the reference-source fixtures
without a GPU.  (The -m gpu suite remains the parity test proper: same comparisons on the device, through the C ABI.)"""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden, rel_elem, rel_max
from oracle import qbold_oracle as o
from oracle import philox

SIG_TOL, GRAD_TOL = 1e-5, 1e-4            # the bars of the GPU parity tests


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope='module')
def emu(tmp_path_factory):
    gxx = shutil.which('g++')
    if not gxx:
        pytest.skip('g++ not available')
    out = str(tmp_path_factory.mktemp('emu') / 'libforward_emu.so')
    subprocess.run([gxx, '-O1', '-std=c++17', '-pthread', '-DQB_HOST_EMU', '-fvisibility=hidden', '-fno-gnu-unique', '-I',
                    os.path.join(ROOT, 'tests', 'host_emu', 'shim'),
                    '-I', os.path.join(ROOT, 'qbold_vi_b200', 'csrc'), '-shared', '-fPIC',
                    os.path.join(ROOT, 'tests', 'host_emu', 'forward_host.cpp'), '-o', out], check=True, capture_output=True,
                   timeout=600)
    return C.CDLL(out)


@pytest.fixture(scope='module')
def qb():
    import qbold_vi_b200 as qb
    if not os.path.exists(qb._lib.LIB_PATH):
        qb.build_library()
    return qb


def _cfg():
    cfg = o.default_config()
    cfg['simulate_noise'] = 'False'
    return cfg


def _pair(emu, layer, x, g, bwd=True, grid=2, block=64):
    n = x.shape[0]
    x = np.ascontiguousarray(x, np.float32)
    g = None if g is None else np.ascontiguousarray(g, np.float32)
    sig = np.full((n, layer.n_tau), np.nan, np.float32)
    grad = np.full((n, 2), np.nan, np.float32)
    emu.qb_emu_forward_pair(C.byref(layer.params), _p(x), _p(g), _p(sig), _p(grad), C.c_int64(n), int(bwd), grid, block)
    return sig, grad


def _generic(emu, params, x, g, n_tau, bwd=True, hct=False, path=0, grid=2, block=64):
    n, w = x.shape
    x = np.ascontiguousarray(x, np.float32)
    g = None if g is None else np.ascontiguousarray(g, np.float32)
    sig = np.full((n, n_tau), np.nan, np.float32)
    grad = np.full((n, w), np.nan, np.float32)
    rc = emu.qb_emu_forward(C.byref(params), _p(x), _p(g), _p(sig), _p(grad), C.c_int64(n), int(bwd), int(hct), path, grid,
                            block)
    assert rc == 0
    return sig, grad


def _voxels(n, seed):
    r = np.random.default_rng(seed)
    return np.stack([r.uniform(0.04, 0.84, n), r.uniform(0.001, 0.201, n)], -1).astype(np.float32)


def test_headline_kernel_on_the_host_matches_oracle_and_reference_fixture(emu, qb):
    """k_forward_pair<BWD> (BASELINE config 2's kernel): random voxels incl. a ragged last pair against the float64
    oracle, and the reference-source fixture (signal, gradient with random and with all-ones upstream)."""
    layer = qb.SignalGenerationLayer(_cfg(), True, True)
    ph = o.parse_params(_cfg())
    for n, grid, block in ((1, 1, 32), (2, 1, 32), (257, 2, 64), (1001, 3, 96)):
        x = _voxels(n, 10 + n)
        g = np.random.default_rng(n).standard_normal((n, 11)).astype(np.float32)
        sig, grad = _pair(emu, layer, x, g, grid=grid, block=block)
        s64, g64 = o.forward_backward(ph, x, g, dtype=np.float64)
        assert rel_elem(sig, s64) < SIG_TOL and rel_elem(sig, s64) < 2e-6            # measured 5e-7
        assert rel_max(grad, g64) < GRAD_TOL and rel_max(grad, g64) < 2e-6           # measured 1.2e-7
        fwd_only, untouched = _pair(emu, layer, x, None, bwd=False, grid=grid, block=block)
        assert np.array_equal(fwd_only, sig) and np.isnan(untouched).all()           # K1 == the value half of K1b
    d = golden('ref_shim_forward.npz')
    sig, grad = _pair(emu, layer, d['oef_dbv'], d['g_rand'])
    assert rel_elem(sig, d['signal_f1_b1']) < SIG_TOL
    assert rel_max(grad, d['grad_rand_f1_b1']) < GRAD_TOL
    _, grad1 = _pair(emu, layer, d['oef_dbv'], None)                                 # g_signal NULL == all ones
    assert rel_max(grad1, d['grad_ones_f1_b1']) < GRAD_TOL


def test_results_do_not_depend_on_the_launch_shape(emu, qb):
    """Voxels come from a work counter and warps pair them up two at a time: any grid / block split gives the same bits."""
    layer = qb.SignalGenerationLayer(_cfg(), True, True)
    x, g = _voxels(131, 3), np.random.default_rng(3).standard_normal((131, 11)).astype(np.float32)
    ref = _pair(emu, layer, x, g, grid=1, block=32)
    for grid, block in ((1, 256), (4, 64), (7, 32)):
        got = _pair(emu, layer, x, g, grid=grid, block=block)
        assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])


@pytest.mark.parametrize('full,blood', [(True, True), (True, False), (False, True), (False, False)])
def test_one_voxel_per_warp_kernel_all_model_variants(emu, qb, full, blood):
    """k_forward<BWD, HCT, PATH>: the scheduled path and the column-group path (schedule switched off in the parameter
    block) for the four full-model / blood combinations, against the reference-source fixture; variable Hct with the
    3-column gradient against its fixture."""
    key = 'f%d_b%d' % (int(full), int(blood))
    d = golden('ref_shim_forward.npz')
    layer = qb.SignalGenerationLayer(_cfg(), full, blood)
    plain = type(layer.params)()
    C.memmove(C.byref(plain), C.byref(layer.params), C.sizeof(plain))
    plain.sched_phases = 0                                                           # -> kCols
    for params, path in ((layer.params, 0), (plain, 1)):
        if path == 0 and layer.params.sched_phases == 0:
            continue
        sig, grad = _generic(emu, params, d['oef_dbv'], d['g_rand'], 11, path=path)
        assert rel_elem(sig, d['signal_' + key]) < SIG_TOL
        assert rel_max(grad, d['grad_rand_' + key]) < GRAD_TOL
    h = golden('ref_shim_forward_hct.npz')
    lay_h = qb.SignalGenerationLayer(_cfg(), full, blood, variable_hct=True)
    sig, grad = _generic(emu, lay_h.params, h['oef_dbv_hct'], h['g_rand'], 11, hct=True,
                         path=0 if lay_h.params.sched_phases > 0 else 1)
    assert rel_elem(sig, h['signal_' + key]) < SIG_TOL
    for j in range(3):
        assert rel_max(grad[:, j], h['grad_rand_' + key][:, j]) < GRAD_TOL


@pytest.mark.parametrize('blood', [True, False])
def test_log_linear_kernel(emu, qb, blood):
    """k_loglinear (full_model = False, the HBM-bound branch): tiles through dynamic shared memory, 16-byte and ragged
    stores; against the reference-source fixture and, on a ragged batch, the float64 oracle."""
    key = 'f0_b%d' % int(blood)
    d = golden('ref_shim_forward.npz')
    layer = qb.SignalGenerationLayer(_cfg(), False, blood)
    ph = o.parse_params(_cfg())

    def run(x, g, bwd=True):
        n = x.shape[0]
        x = np.ascontiguousarray(x, np.float32)
        g = None if g is None else np.ascontiguousarray(g, np.float32)
        sig, grad = np.full((n, 11), np.nan, np.float32), np.full((n, 2), np.nan, np.float32)
        emu.qb_emu_loglinear(C.byref(layer.params), _p(x), _p(g), _p(sig), _p(grad), C.c_int64(n), int(bwd))
        return sig, grad

    sig, grad = run(d['oef_dbv'], d['g_rand'])
    assert rel_elem(sig, d['signal_' + key]) < SIG_TOL and rel_max(grad, d['grad_rand_' + key]) < GRAD_TOL
    _, grad1 = run(d['oef_dbv'], None)
    assert rel_max(grad1, d['grad_ones_' + key]) < GRAD_TOL
    x = _voxels(600 + 3, 21)                                       # three CTAs, the last one ragged (count % 4 != 0)
    g = np.random.default_rng(21).standard_normal((603, 11)).astype(np.float32)
    sig, grad = run(x, g)
    s64, g64 = o.forward_backward(ph, x, g, False, blood, np.float64)
    assert rel_elem(sig, s64) < SIG_TOL and rel_max(grad, g64) < GRAD_TOL
    assert np.array_equal(run(x, None, bwd=False)[0], sig)


def test_24_tau_grid_on_the_multi_group_path(emu, qb):
    """16 distinct |tau| columns (the 24-tau protocol): column groups beyond the first are read from the parameter block."""
    cfg = dict(_cfg(), tau_start='-0.028', tau_end='0.065', tau_step='0.004')
    layer = qb.SignalGenerationLayer(cfg, True, True)
    assert layer.n_tau == 24 and layer.params.n_cols > 8
    ph = o.parse_params(cfg)
    x = _voxels(96, 5)
    g = np.random.default_rng(5).standard_normal((96, 24)).astype(np.float32)
    sig, grad = _generic(emu, layer.params, x, g, 24, path=2)
    s64, g64 = o.forward_backward(ph, x, g, dtype=np.float64)
    assert rel_elem(sig, s64) < SIG_TOL and rel_max(grad, g64) < GRAD_TOL


def test_misalignment_kernel_with_recorded_and_philox_draws(emu, qb):
    """k_misalign: the reference's recorded draws (variable-Hct fixture) and the in-kernel Philox path against
    oracle/philox.py draws for the same (seed, global voxel index)."""
    h = golden('ref_shim_forward_hct.npz')
    lay = qb.SignalGenerationLayer(_cfg(), True, True, variable_hct=True)
    x = np.ascontiguousarray(h['oef_dbv_hct'], np.float32)
    sig, _ = _generic(emu, lay.params, x, None, 11, bwd=True, hct=True, path=0)
    u, idx, eps = (np.ascontiguousarray(h[k], t) for k, t in (('mis_u01', np.float32), ('mis_index', np.int32),
                                                              ('mis_eps', np.float32)))
    emu.qb_emu_misalign(C.byref(lay.params), _p(x), C.c_int64(x.shape[0]), C.c_float(float(h['mis_prob'])), _p(u), _p(idx),
                        _p(eps), C.c_uint64(0), C.c_uint64(0), _p(sig), 1, 2, 64)
    assert rel_elem(sig, h['signal_misaligned']) < SIG_TOL
    # Philox path, fixed Hct: the oracle is fed the draws oracle/philox.py derives for (seed, offset + voxel)
    layer = qb.SignalGenerationLayer(_cfg(), True, True)
    ph = o.parse_params(_cfg())
    n, seed, offset, prob = 300, 0x1234567890ABCDEF, (1 << 33) + 11, 0.35
    xv = _voxels(n, 8)
    clean, _ = _pair(emu, layer, xv, None, bwd=False)
    got = clean.copy()
    emu.qb_emu_misalign(C.byref(layer.params), _p(xv), C.c_int64(n), C.c_float(prob), None, None, None, C.c_uint64(seed),
                        C.c_uint64(offset), _p(got), 0, 2, 64)
    u, idx, eps = philox.misalign_draws(seed, offset + np.arange(n, dtype=np.uint64), 11)
    want = o.forward_misaligned(ph, xv, prob, u, idx, eps, dtype=np.float64)
    assert rel_elem(got, want) < SIG_TOL
    hit = u < prob
    assert 0.2 < hit.mean() < 0.5 and np.array_equal(got[~hit], clean[~hit]) and np.array_equal(got[:, :5], clean[:, :5])


# ------------------------------------------------------------------------------------------- likelihood-side kernels (K2)
@pytest.fixture(scope='module')
def emu_elbo(tmp_path_factory):
    gxx = shutil.which('g++')
    if not gxx:
        pytest.skip('g++ not available')
    out = str(tmp_path_factory.mktemp('emu_elbo') / 'libelbo_emu.so')
    subprocess.run([gxx, '-O1', '-std=c++17', '-pthread', '-DQB_HOST_EMU', '-fvisibility=hidden', '-fno-gnu-unique', '-I',
                    os.path.join(ROOT, 'tests', 'host_emu', 'shim'),
                    '-I', os.path.join(ROOT, 'qbold_vi_b200', 'csrc'), '-shared', '-fPIC',
                    os.path.join(ROOT, 'tests', 'host_emu', 'elbo_host.cpp'), '-o', out], check=True, capture_output=True,
                   timeout=900)
    return C.CDLL(out)


def _trainer(qb, **kw):
    args = dict(student_t_df=200, multi_image_normalisation=False, use_mvg=True, use_population_prior=False,
                predict_log_data=False, seed=1234)
    args.update(kw)
    return qb.EncoderTrainer(_cfg(), **args)


def _elbo(lib, params, q, sigma, y, mask, prior, eps=None, eps_kl=None, kl_samples=70, seed=0, offset=0, pair=1, path=0,
          grid=2, seed_dev=None, kl_weight=1.0, mask_sum=None, inv_dev=False):
    n, nt = q.shape[0], sigma.shape[1]
    arrs = [None if a is None else np.ascontiguousarray(a, np.float32) for a in (q, sigma, y, mask, prior, eps, eps_kl)]
    gq, gs = np.full((n, 5), np.nan, np.float32), np.full((n, nt), np.nan, np.float32)
    nm, km, sums = np.full(n, np.nan, np.float32), np.full(n, np.nan, np.float32), np.zeros(4, np.float64)
    inv = np.float32(1.0 / float(mask.sum() if mask_sum is None else mask_sum))
    inv_arr = np.array([inv], np.float32)
    sd = None if seed_dev is None else np.array([seed_dev], np.uint64)
    rc = lib.qb_emu_elbo(C.byref(params), *[_p(a) for a in arrs], C.c_uint64(seed), _p(sd), C.c_uint64(offset), kl_samples,
                         C.c_float(0.0 if inv_dev else inv), _p(inv_arr) if inv_dev else None, C.c_float(kl_weight),
                         C.c_int64(n), _p(gq), _p(gs), _p(nm), _p(km), _p(sums), pair, path, grid)
    assert rc == 0
    return dict(grad_q=gq, grad_sigma=gs, nll_map=nm, kl_map=km, nll=sums[0] * float(inv), kl=sums[1] * float(inv),
                mask_sum=sums[2], non_finite=sums[3])


def _elbo_batch(ph, n, seed):
    r = np.random.default_rng(seed)
    q = np.stack([r.normal(-0.3, 0.7, n), r.normal(0, 0.6, n), r.normal(-1.2, 0.7, n), r.normal(0, 0.6, n),
                  r.normal(0, 0.8, n)], -1).astype(np.float32)
    prior = (q + r.normal(0, 0.3, (n, 5))).astype(np.float32)
    sigma = np.exp(r.normal(np.log(0.05), 0.2, (n, 11))).astype(np.float32)
    truth = np.stack([r.uniform(0.1, 0.7, n), r.uniform(0.005, 0.15, n)], -1)
    data = (o.forward(ph, truth, dtype=np.float64) * 100 * (1 + 0.02 * r.standard_normal((n, 11)))).astype(np.float32)
    mask = (r.uniform(size=n) > 0.3).astype(np.float32)
    return q, prior, sigma, data * mask[:, None], mask


@pytest.mark.parametrize('tag', ['optimal', 'multinorm', 'studentt'])
def test_fused_elbo_kernels_on_the_host_match_the_reference_fixture(emu_elbo, qb, tag):
    """k_elbo_pair<HAS_PRIOR> (production) and the generic k_elbo on the scheduled and the column-group path: loss terms,
    maps and both gradients against the fixture recorded from the reference's model.py + signals.py."""
    e = golden('ref_shim_elbo_%s.npz' % tag)
    tr = _trainer(qb, student_t_df=float(e['student_t_df']), multi_image_normalisation=bool(e['multi_image_normalisation']))
    layer = qb.SignalGenerationLayer(_cfg(), True, True)
    params = tr._params_for(layer)
    plain = type(params)()
    C.memmove(C.byref(plain), C.byref(params), C.sizeof(plain))
    plain.sched_phases = 0
    for P, pair, path in ((params, 1, 0), (params, 0, 0), (plain, 0, 1)):
        r = _elbo(emu_elbo, P, e['q'], e['sigma'], e['data'], e['mask'], e['prior'], e['eps'], e['eps_kl'], pair=pair, path=path)
        assert rel_elem(r['nll'], e['nll']) < GRAD_TOL and rel_elem(r['kl'], e['kl']) < GRAD_TOL
        assert rel_max(r['nll_map'], e['nll_map']) < GRAD_TOL
        assert rel_max(r['grad_q'], e['grad_q_nll'] + e['grad_q_kl']) < GRAD_TOL
        assert rel_max(r['grad_sigma'], e['grad_sigma']) < GRAD_TOL
        assert r['mask_sum'] == e['mask'].sum() and r['non_finite'] == 0
        dead = e['mask'] == 0
        assert np.all(r['grad_q'][dead] == 0) and np.all(r['grad_sigma'][dead] == 0)     # model.py:564,661
    # no prior: likelihood term only
    r = _elbo(emu_elbo, params, e['q'], e['sigma'], e['data'], e['mask'], None, e['eps'], None, kl_samples=0)
    assert rel_elem(r['nll'], e['nll']) < GRAD_TOL and r['kl'] == 0.0
    assert rel_max(r['grad_q'], e['grad_q_nll']) < GRAD_TOL


def test_in_kernel_philox_draws_and_the_device_resident_key(emu_elbo, qb):
    """No explicit draws: in-kernel Philox + Box-Muller against the oracle fed oracle/philox.py draws for the same
    (seed, global voxel index).  The Philox key read through a pointer (qbold_elbo_fused_graph, the captured training
    step) and 1/sum(mask) read through a pointer give the same bits as the by-value launch; so does any launch shape."""
    ph = o.parse_params(_cfg())
    n, S, seed, off = 384, 70, 0x9E3779B97F4A7C15 + 99, (1 << 32) + 1_234_567
    q, prior, sigma, data, mask = _elbo_batch(ph, n, 31)
    idx = np.arange(n, dtype=np.uint64) + np.uint64(off)
    eps, eps_kl = philox.reparam_eps(seed, idx), philox.kl_eps(seed, idx, S)
    ref = o.elbo_and_grads(ph, q, sigma, data, mask, prior, eps, eps_kl, np.float64)
    layer = qb.SignalGenerationLayer(_cfg(), True, True)
    params = _trainer(qb)._params_for(layer)
    r = _elbo(emu_elbo, params, q, sigma, data, mask, prior, kl_samples=S, seed=seed, offset=off)
    assert rel_elem(r['nll'], ref['nll']) < GRAD_TOL and rel_elem(r['kl'], ref['kl']) < GRAD_TOL
    assert rel_max(r['nll_map'], ref['nll_map']) < GRAD_TOL and rel_max(r['kl_map'], ref['kl_map']) < GRAD_TOL
    assert rel_max(r['grad_q'], ref['grad_q']) < GRAD_TOL and rel_max(r['grad_sigma'], ref['grad_sigma']) < GRAD_TOL
    by_ptr = _elbo(emu_elbo, params, q, sigma, data, mask, prior, kl_samples=S, seed=12345, seed_dev=seed, offset=off,
                   inv_dev=True, grid=3)
    for k in ('grad_q', 'grad_sigma', 'nll_map', 'kl_map'):
        assert np.array_equal(by_ptr[k], r[k]), k
    # explicit draws equal to the Philox draws reproduce the likelihood side bit for bit; the KL differs only by the
    # SFU-approximation Box-Muller of the sampling loop (emulated with the exact functions here)
    ex = _elbo(emu_elbo, params, q, sigma, data, mask, prior, eps, eps_kl, kl_samples=S)
    assert rel_max(ex['nll_map'], r['nll_map']) < 1e-6 and rel_max(ex['kl_map'], r['kl_map']) < 1e-4
    # shards of the batch with the global mask count and the shard's global offset: same per-voxel results
    h = 192
    parts = [_elbo(emu_elbo, params, q[s], sigma[s], data[s], mask[s], prior[s], kl_samples=S, seed=seed,
                   offset=off + s.start, mask_sum=mask.sum()) for s in (slice(0, h), slice(h, n))]
    assert np.array_equal(np.concatenate([p['grad_q'] for p in parts]), r['grad_q'])
    assert abs(parts[0]['nll'] + parts[1]['nll'] - r['nll']) < 1e-6 * abs(r['nll'])


def test_kl_reparam_and_posterior_statistics_kernels(emu_elbo, qb):
    """k_kl (70-sample Monte Carlo with recorded draws, and the closed form), k_reparam and k_posterior_stats against
    the reference-source fixtures and the oracle."""
    e = golden('ref_shim_elbo_optimal.npz')
    n = e['q'].shape[0]
    q, prior, mask = (np.ascontiguousarray(e[k], np.float32) for k in ('q', 'prior', 'mask'))
    eps_kl = np.ascontiguousarray(e['eps_kl'], np.float32)
    kl_map, grad = np.full(n, np.nan, np.float32), np.full((n, 5), np.nan, np.float32)
    emu_elbo.qb_emu_kl(_p(q), _p(prior), _p(mask), _p(eps_kl), C.c_uint64(0), C.c_uint64(0), 70, C.c_int64(n), _p(kl_map),
                       _p(grad), 1)
    assert rel_elem(kl_map.sum() / mask.sum(), e['kl']) < GRAD_TOL
    assert rel_max(grad / mask.sum(), e['grad_q_kl']) < GRAD_TOL
    emu_elbo.qb_emu_kl(_p(q), _p(prior), _p(mask), None, C.c_uint64(0), C.c_uint64(0), 0, C.c_int64(n), _p(kl_map), _p(grad), 1)
    assert rel_max(kl_map, np.where(mask > 0, o.closed_form_kl(prior, q), 0)) < 1e-5
    # reparameterised sample (model.py:21-50)
    eps = np.ascontiguousarray(e['eps'], np.float32)
    out = np.full((n, 2), np.nan, np.float32)
    emu_elbo.qb_emu_reparam(_p(q), _p(eps), C.c_uint64(0), C.c_uint64(0), C.c_int64(n), _p(out))
    want, _ = o.reparam_sample(q, eps, True, np.float64)
    assert rel_max(out, want) < 1e-6
    seed, off = 77, 5000
    emu_elbo.qb_emu_reparam(_p(q), None, C.c_uint64(seed), C.c_uint64(off), C.c_int64(n), _p(out))
    want, _ = o.reparam_sample(q, philox.reparam_eps(seed, np.arange(n, dtype=np.uint64) + np.uint64(off)), True, np.float64)
    assert rel_max(out, want) < 1e-5
    # posterior statistics (model.py:326-343): in-kernel draws, 64 samples per voxel, against explicit oracle samples
    layer = qb.SignalGenerationLayer(_cfg(), True, True)
    S = 64
    mean3, var3 = np.full((n, 3), np.nan, np.float32), np.full((n, 3), np.nan, np.float32)
    emu_elbo.qb_emu_posterior_stats(C.c_float(layer.params.dw_k), _p(q), None, C.c_uint64(seed), C.c_uint64(off), S,
                                    C.c_int64(n), _p(mean3), _p(var3), 1)
    draws = philox.kl_eps(seed, np.arange(n, dtype=np.uint64) + np.uint64(off), S)               # [n, S, 2]
    smp = np.stack([o.reparam_sample(q, draws[:, s], True, np.float64)[0] for s in range(S)], 1)  # [n, S, 2]
    r2p = float(layer.params.dw_k) * smp[..., 0] * smp[..., 1]
    vals = np.concatenate([smp, r2p[..., None]], -1)
    assert rel_max(mean3, vals.mean(1)) < 1e-4 and rel_max(var3, vals.var(1)) < 2e-3


# ------------------------------------------------------------------------------------------- streaming generation (K3)
@pytest.fixture(scope='module')
def emu_gen(tmp_path_factory):
    gxx = shutil.which('g++')
    if not gxx:
        pytest.skip('g++ not available')
    out = str(tmp_path_factory.mktemp('emu_gen') / 'libgen_emu.so')
    subprocess.run([gxx, '-O1', '-std=c++17', '-pthread', '-DQB_HOST_EMU', '-fvisibility=hidden', '-fno-gnu-unique', '-I',
                    os.path.join(ROOT, 'tests', 'host_emu', 'shim'),
                    '-I', os.path.join(ROOT, 'qbold_vi_b200', 'csrc'), '-shared', '-fPIC',
                    os.path.join(ROOT, 'tests', 'host_emu', 'generate_host.cpp'), '-o', out], check=True, capture_output=True,
                   timeout=900)
    return C.CDLL(out)


def _generate(lib, layer, oefs, dbvs, perm, seed, n_chunks, snr_u01=None, noise_eps=None, pair=1, path=0):
    """generate_from_marginals (qbold_vi_b200/signals.py) on the emulated kernels."""
    oefs, dbvs = np.ascontiguousarray(oefs, np.float32), np.ascontiguousarray(dbvs, np.float32)
    perm = None if perm is None else np.ascontiguousarray(perm, np.int64)
    nt, total = layer.n_tau, oefs.size * dbvs.size
    chunk = total // n_chunks
    n_x = chunk * n_chunks
    x = np.full((n_x, nt), np.nan, np.float32)
    y = np.full((total, 3), np.nan, np.float32)
    rc = lib.qb_emu_generate(C.byref(layer.params), _p(oefs), C.c_int64(oefs.size), _p(dbvs), C.c_int64(dbvs.size), _p(perm),
                             C.c_uint64(seed), C.c_int64(0), C.c_int64(n_x), _p(x), _p(y), pair, path, 2)
    assert rc == 0
    if total > n_x:
        tail = np.full((total - n_x, 3), np.nan, np.float32)
        lib.qb_emu_generate(C.byref(layer.params), _p(oefs), C.c_int64(oefs.size), _p(dbvs), C.c_int64(dbvs.size), _p(perm),
                            C.c_uint64(seed), C.c_int64(n_x), C.c_int64(total - n_x), None, _p(tail), pair, path, 1)
        y[n_x:] = tail
    clean = x.copy()
    if snr_u01 is not None or noise_eps is None:
        s = None if snr_u01 is None else np.ascontiguousarray(snr_u01[:n_x], np.float32)
        e = None if noise_eps is None else np.ascontiguousarray(noise_eps[:n_x], np.float32)
        lib.qb_emu_add_noise_chunked(C.byref(layer.params), _p(x), C.c_int64(chunk), n_chunks, _p(s), _p(e),
                                     C.c_uint64((seed ^ 0x5DEECE66D) & 0xFFFFFFFFFFFFFFFF), C.c_uint64(0), 2)
    return x, y, clean


@pytest.mark.parametrize('tag', ['u10', 'u0'])
def test_dataset_generation_kernels_match_the_reference_source(emu_gen, qb, tag):
    """k_generate_pair / k_generate + the chunked noise kernels against create_synthetic_dataset of the reference
    (recorded permutation and draws; the S^2 % 10 trailing rows are dropped from x only, signals.py:283-287)."""
    d = golden('ref_shim_dataset_%s.npz' % tag)
    ty = d['train_y']
    grid = np.empty((529, 2), np.float32)
    grid[d['perm']] = ty[:, :2]
    oefs, dbvs = grid.reshape(23, 23, 2)[:, 0, 0].copy(), grid.reshape(23, 23, 2)[0, :, 1].copy()
    layer = qb.SignalGenerationLayer(o.default_config(), True, True)                       # noise ON, as in the INI
    for pair, path in ((1, 0), (0, 0)):
        x, y, _ = _generate(emu_gen, layer, oefs, dbvs, d['perm'], 0, 10, d['snr_u01'].reshape(-1), d['noise_eps'],
                            pair=pair, path=path)
        assert x.shape == (520, 11) and y.shape == (529, 3)
        assert rel_elem(y, ty) < 1e-6
        assert rel_elem(x, d['train_x']) < 2 * SIG_TOL


def test_dataset_generation_with_the_feistel_shuffle_and_philox_noise(emu_gen, qb):
    """No recorded draws: the keyed Feistel shuffle and the Philox noise streams, against the oracle fed
    oracle/philox.py's permutation and draws."""
    cfg = o.default_config()
    ph = o.parse_params(cfg)
    layer = qb.SignalGenerationLayer(cfg, True, True, seed=4242)
    r = np.random.default_rng(8)
    oefs = r.uniform(0.05, 0.8, 40).astype(np.float32)
    dbvs = r.uniform(0.003, 0.195, 30).astype(np.float32)
    x, y, clean = _generate(emu_gen, layer, oefs, dbvs, None, 4242, 10)
    perm = philox.feistel_permute(np.arange(1200), 1200, 4242)
    assert sorted(perm.tolist()) == list(range(1200))
    snr = np.concatenate([philox.snr_u01(4242 ^ 0x5DEECE66D, np.arange(i * 120, (i + 1) * 120)) for i in range(10)])
    eps = np.concatenate([philox.noise_eps(4242 ^ 0x5DEECE66D, np.arange(i * 120, (i + 1) * 120), 11) for i in range(10)])
    xo, yo = o.synthetic_dataset_from_draws(ph, oefs, dbvs, perm, snr * np.float32(70) + np.float32(50), eps)
    assert rel_elem(y, yo) < 1e-6
    assert rel_elem(clean, o.forward(ph, yo[:, :2], dtype=np.float64)) < SIG_TOL
    assert rel_elem(x, xo) < 2 * SIG_TOL


# ------------------------------------------------------------------------------------------- losses either side (8f-1/2)
@pytest.fixture(scope='module')
def emu_losses(tmp_path_factory):
    gxx = shutil.which('g++')
    if not gxx:
        pytest.skip('g++ not available')
    out = str(tmp_path_factory.mktemp('emu_losses') / 'liblosses_emu.so')
    subprocess.run([gxx, '-O1', '-std=c++17', '-pthread', '-DQB_HOST_EMU', '-fvisibility=hidden', '-fno-gnu-unique', '-I',
                    os.path.join(ROOT, 'tests', 'host_emu', 'shim'),
                    '-I', os.path.join(ROOT, 'qbold_vi_b200', 'csrc'), '-shared', '-fPIC',
                    os.path.join(ROOT, 'tests', 'host_emu', 'losses_host.cpp'), '-o', out], check=True, capture_output=True,
                   timeout=900)
    return C.CDLL(out)


def test_tv_pretraining_and_population_kl_kernels_match_the_reference_source(emu_losses):
    """k_smoothness, k_synth_nll (mvg / diagonal, fixed and learned InverseGamma), k_diag_kl, k_mog_kl: value and gradient
    against the fixture recorded from the reference's model.py."""
    a = golden('ref_shim_adjacent.npz')
    shp = tuple(int(v) for v in a['shape'])
    n = int(np.prod(shp))
    mask = np.ascontiguousarray(a['mask'], np.float32)
    msum = float(mask.sum())
    # total variation (model.py:726-754): x / y neighbours inside the mask
    for tag, c in (('mvg', 5), ('diag', 4)):
        q = np.ascontiguousarray(a['q5'][:, :c], np.float32)
        tv, grad = np.zeros(1, np.float64), np.full((n, c), np.nan, np.float32)
        emu_losses.qb_emu_smoothness(_p(q), c, _p(mask), C.c_int64(shp[0]), shp[1], shp[2], shp[3], C.c_float(1.0 / msum),
                                     None, _p(tv), _p(grad), 2)
        assert rel_elem(tv[0] / msum, a['tv_' + tag]) < GRAD_TOL
        assert rel_max(grad, a['tv_%s_grad' % tag]) < GRAD_TOL
        scale = np.array([1.0 / msum], np.float32)                                       # the same through a pointer
        tv2, grad2 = np.zeros(1, np.float64), np.full((n, c), np.nan, np.float32)
        emu_losses.qb_emu_smoothness(_p(q), c, _p(mask), C.c_int64(shp[0]), shp[1], shp[2], shp[3], C.c_float(0.0),
                                     _p(scale), _p(tv2), _p(grad2), 1)
        assert np.array_equal(grad2, grad) and abs(tv2[0] - tv[0]) < 1e-9 * abs(tv[0])
    # pre-training NLL (model.py:449-514)
    for tag, use_mvg, ig in (('mvg', 1, (0.0, 0.0)), ('mvg_ig', 1, (3.0, 0.15)), ('diag', 0, (0.0, 0.0)),
                             ('diag_ig', 0, (3.0, 0.15))):
        c = 5 if use_mvg else 4
        pred = np.ascontiguousarray(a['q5'][:, :c], np.float32)
        labels = np.ascontiguousarray(a['synth_%s_labels' % tag], np.float32)
        total, grad = np.zeros(1, np.float64), np.full((n, c), np.nan, np.float32)
        emu_losses.qb_emu_synth_nll(_p(labels), 3, _p(pred), c, use_mvg, C.c_double(ig[0]), C.c_double(ig[1]), None,
                                    C.c_int64(n), C.c_float(1.0 / n), None, _p(grad), _p(total), None, 1)
        assert rel_elem(total[0] / n, a['synth_' + tag]) < GRAD_TOL
        assert rel_max(grad, a['synth_%s_grad' % tag]) < GRAD_TOL
    # infer_inv_gamma (model.py:454-455,493-496): the 4 hyper-prior channels of voxel 0 (:494), read through a pointer
    pred8 = np.ascontiguousarray(a['synth_iginf_pred'], np.float32)
    labels = np.ascontiguousarray(a['synth_iginf_labels'], np.float32)
    ig4 = np.ascontiguousarray(pred8[0, 4:8])
    total, grad, igs = np.zeros(1, np.float64), np.full((n, 4), np.nan, np.float32), np.zeros(4, np.float64)
    emu_losses.qb_emu_synth_nll(_p(labels), 3, _p(pred8), 8, 0, C.c_double(0.0), C.c_double(0.0), _p(ig4), C.c_int64(n),
                                C.c_float(1.0 / n), None, _p(grad), _p(total), _p(igs), 2)
    assert rel_elem(total[0] / n, a['synth_iginf']) < GRAD_TOL
    assert rel_max(grad, a['synth_iginf_grad'][:, :4]) < GRAD_TOL
    # diagonal KL (model.py:685-708) against a fixed prior
    q4, p4 = np.ascontiguousarray(a['q5'][:, :4], np.float32), np.ascontiguousarray(a['prior5'][:, :4], np.float32)
    kl, grad = np.full(n, np.nan, np.float32), np.full((n, 4), np.nan, np.float32)
    emu_losses.qb_emu_diag_kl(_p(q4), 4, _p(p4), 4, _p(mask), C.c_int64(n), _p(kl), _p(grad), 4, None, 0, 1)
    assert rel_elem(kl.sum() / msum, a['kl_diag']) < GRAD_TOL
    assert rel_max(grad / msum, a['kl_diag_grad']) < GRAD_TOL
    ref, _, _ = o.diag_kl(a['prior5'][:, :4], a['q5'][:, :4])
    assert rel_max(kl, ref * (mask > 0)) < 1e-5
    # mixture-of-Gaussians population prior (model.py:666-684), the reference's recorded draw per voxel
    pred16 = np.ascontiguousarray(a['kl_mog_pred'], np.float32)
    eps = np.ascontiguousarray(a['kl_mog_eps'], np.float32)
    kl, grad = np.full(n, np.nan, np.float32), np.full((n, 16), np.nan, np.float32)
    emu_losses.qb_emu_mog_kl(_p(pred16), 3, _p(mask), _p(eps), C.c_uint64(0), C.c_uint64(0), C.c_int64(n), _p(kl), _p(grad), 2)
    assert rel_elem(kl.sum() / msum, a['kl_mog']) < GRAD_TOL
    assert rel_max(grad / msum, a['kl_mog_grad']) < GRAD_TOL


@pytest.mark.parametrize('tag', ['optimal', 'multinorm', 'studentt'])
def test_standalone_likelihood_and_likelihood_map_kernels(emu_elbo, qb, tag):
    """k_nll (fine_tune_loss_fn on predictions that already exist: value, d/dsigma) against the reference fixture, and
    k_nll_map_pair / k_nll_map (likelihood map of save_predictions, model.py:808-817: one forward pass per sample)
    against the oracle, with explicit draws and with in-kernel Philox draws."""
    e = golden('ref_shim_elbo_%s.npz' % tag)
    tr = _trainer(qb, student_t_df=float(e['student_t_df']), multi_image_normalisation=bool(e['multi_image_normalisation']))
    layer = qb.SignalGenerationLayer(_cfg(), True, True)
    params = tr._params_for(layer)
    n = e['q'].shape[0]
    y, pred, sigma, mask = (np.ascontiguousarray(e[k], np.float32) for k in ('data', 'pred', 'sigma', 'mask'))
    nll, d_pred, d_sigma = np.full(n, np.nan, np.float32), np.full((n, 11), np.nan, np.float32), np.full((n, 11), np.nan, np.float32)
    emu_elbo.qb_emu_nll(C.byref(params), _p(y), _p(pred), _p(sigma), _p(mask), C.c_int64(n), _p(nll), _p(d_pred), _p(d_sigma), 1)
    assert rel_max(nll, e['nll_map']) < GRAD_TOL and rel_elem(nll.sum() / mask.sum(), e['nll']) < GRAD_TOL
    assert rel_max(d_sigma / mask.sum(), e['grad_sigma']) < GRAD_TOL
    if tag == 'studentt':
        return                                              # the oracle's per-sample likelihood covers the Gaussian case
    ph = o.parse_params(_cfg())
    q = np.ascontiguousarray(e['q'], np.float32)
    S, se = 6, 2
    eps = np.random.default_rng(4).standard_normal((n, S, 2)).astype(np.float32)
    yt = np.concatenate([e['data'], e['mask'][:, None]], -1)

    def oracle_map(draws):
        ref = np.zeros(n)
        for s in range(S):
            smp, _ = o.reparam_sample(e['q'], draws[:, s], True, np.float64)
            ref += o.fine_tune_nll(yt, o.forward(ph, smp, dtype=np.float64), e['sigma'], se, np.float64, return_mean=False,
                                   multi_image_normalisation=bool(e['multi_image_normalisation'])).reshape(-1)
        return ref / S

    for pair, path in ((1, 0), (0, 0)):
        got = np.full(n, np.nan, np.float32)
        rc = emu_elbo.qb_emu_nll_map(C.byref(params), _p(q), _p(sigma), _p(y), _p(mask), _p(eps), C.c_uint64(0), C.c_uint64(0),
                                     S, C.c_int64(n), _p(got), pair, path, 1)
        assert rc == 0 and rel_max(got, oracle_map(eps)) < GRAD_TOL
    seed, off = 31, 1 << 20
    got = np.full(n, np.nan, np.float32)
    emu_elbo.qb_emu_nll_map(C.byref(params), _p(q), _p(sigma), _p(y), _p(mask), None, C.c_uint64(seed), C.c_uint64(off), S,
                            C.c_int64(n), _p(got), 1, 0, 2)
    draws = philox.kl_eps(seed, np.arange(n, dtype=np.uint64) + np.uint64(off), S)
    assert rel_max(got, oracle_map(draws)) < GRAD_TOL


# ------------------------------------------------------------------------------------------- encoder block streaming (8f-3)
@pytest.fixture(scope='module')
def emu_enc(tmp_path_factory):
    gxx = shutil.which('g++')
    if not gxx:
        pytest.skip('g++ not available')
    out = str(tmp_path_factory.mktemp('emu_enc') / 'libenc_emu.so')
    subprocess.run([gxx, '-O1', '-std=c++17', '-pthread', '-DQB_HOST_EMU', '-fvisibility=hidden', '-fno-gnu-unique', '-I',
                    os.path.join(ROOT, 'tests', 'host_emu', 'shim'),
                    '-I', os.path.join(ROOT, 'qbold_vi_b200', 'csrc'), '-shared', '-fPIC',
                    os.path.join(ROOT, 'tests', 'host_emu', 'encoder_block_host.cpp'), '-o', out], check=True,
                   capture_output=True, timeout=900)
    return C.CDLL(out)


def test_encoder_block_streaming_kernels(emu_enc):
    """The elementwise kernels of the fused encoder block (create_block, model.py:142-174) against the formulas in
    float64 NumPy: gated mix forward (+ the ReLU copy) and backward (incl. the stream-1 addend and the ReLU' of the skip
    branch), ReLU' + bias-gradient column sums (two-stage fixed-order reduction, accumulate flag), and normalise_data
    (model.py:97-113) fused with the move to the z-outer layout (ragged tiles, single- and multi-image reference)."""
    r = np.random.default_rng(12)
    n, c = 777, 60
    skip, r0, z, go, add = (r.standard_normal((n, c)).astype(np.float32) for _ in range(5))
    skip = np.maximum(skip, 0)                                                    # the skip branch ends in a ReLU
    b_r = r.standard_normal(c).astype(np.float32)
    off = np.float32(-1.0)
    out, out_relu = np.full((n, c), np.nan, np.float32), np.full((n, c), np.nan, np.float32)
    emu_enc.qb_emu_block_mix_forward(_p(skip), _p(r0), _p(b_r), _p(z), C.c_float(off), C.c_int64(n), c, _p(out), _p(out_relu), 3)
    g = 1.0 / (1.0 + np.exp(-(z.astype(np.float64) + float(off))))
    rr = r0.astype(np.float64) + b_r
    want = skip * (1 - g) + rr * g
    assert rel_max(out, want) < 1e-6 and np.array_equal(out_relu, np.maximum(out, 0))
    d_skip, d_r, d_z = (np.full((n, c), np.nan, np.float32) for _ in range(3))
    emu_enc.qb_emu_block_mix_backward(_p(go), _p(skip), _p(r0), _p(b_r), _p(z), C.c_float(off), C.c_int64(n), c, 1, _p(add),
                                      _p(d_skip), _p(d_r), _p(d_z), 2)
    assert rel_max(d_skip, (go * (1 - g) + add) * (skip > 0)) < 1e-6
    assert rel_max(d_r, go * g) < 1e-6 and rel_max(d_z, go * (rr - skip) * g * (1 - g)) < 1e-5
    emu_enc.qb_emu_block_mix_backward(_p(go), _p(skip), _p(r0), None, _p(z), C.c_float(off), C.c_int64(n), c, 0, None,
                                      _p(d_skip), _p(d_r), _p(d_z), 1)
    assert rel_max(d_skip, go * (1 - g)) < 1e-6 and rel_max(d_z, go * (r0.astype(np.float64) - skip) * g * (1 - g)) < 1e-5
    # ReLU' + column sums
    y = r.standard_normal((n, c)).astype(np.float32)
    masked, colsum = np.full((n, c), np.nan, np.float32), np.full(c, 7.0, np.float32)
    emu_enc.qb_emu_relu_bwd_colsum(_p(go), _p(y), _p(add), C.c_int64(n), c, _p(masked), _p(colsum), 0, 5)
    want = go * (y > 0) + add
    assert np.array_equal(masked, want.astype(np.float32)) and rel_max(colsum, want.astype(np.float64).sum(0)) < 1e-5
    again = colsum.copy()
    emu_enc.qb_emu_relu_bwd_colsum(_p(go), None, None, C.c_int64(n), c, None, _p(again), 1, 4)     # plain sums, accumulated
    assert rel_max(again, colsum + go.astype(np.float64).sum(0)) < 1e-5
    # normalise_data + z-outer transpose
    for (B, X, Y, Z, T, se, multi) in ((2, 3, 37, 5, 11, 2, 1), (1, 2, 33, 40, 11, 2, 0), (1, 1, 4, 3, 24, 7, 1)):
        data = (r.uniform(0.0, 3.0, (B, X, Y, Z, T)) * (r.uniform(size=(B, X, Y, Z, 1)) > 0.1)).astype(np.float32)
        tp = (T + 3) & ~3
        got = np.full((B, Z, X, Y, tp), np.nan, np.float32)
        emu_enc.qb_emu_normalise_zouter(_p(data), C.c_int64(B), X, Y, Z, T, se, multi, _p(got))
        d = np.clip(data.astype(np.float64), 1e-2, 1e8)
        ref = d[..., se - 1:se + 2].mean(-1, keepdims=True) if multi else d[..., se:se + 1]
        want = np.log(d / ref).transpose(0, 3, 1, 2, 4)
        assert rel_max(got[..., :T], want) < 1e-5 and np.all(got[..., T:] == 0)


def test_noise_model_of_one_layer_call(emu_gen, qb):
    """k_column_sum + k_finish_mean + k_add_noise (signals.py:116-128: the noise std is the batch mean of each image
    over the SNR): explicit draws against the oracle, and the Philox streams against oracle/philox.py's draws."""
    cfg = o.default_config()
    ph = o.parse_params(cfg)
    layer = qb.SignalGenerationLayer(cfg, True, True)
    r = np.random.default_rng(6)
    n = 1500
    clean = o.forward(ph, _voxels(n, 6), dtype=np.float64).astype(np.float32)
    u = r.uniform(size=n).astype(np.float32)
    eps = r.standard_normal((n, 11)).astype(np.float32)
    sig, mean = clean.copy(), np.full(11, np.nan, np.float32)
    emu_gen.qb_emu_add_noise(C.byref(layer.params), _p(sig), C.c_int64(n), _p(u), _p(eps), C.c_uint64(0), C.c_uint64(0),
                             _p(mean), 3)
    assert rel_max(mean, clean.astype(np.float64).mean(0)) < 1e-6
    want = o.add_noise(clean, u * np.float32(70) + np.float32(50), eps, np.float64)
    assert rel_elem(sig, want) < SIG_TOL
    seed, off = 0xABCDEF0123, 7_000_000_000
    sig2 = clean.copy()
    emu_gen.qb_emu_add_noise(C.byref(layer.params), _p(sig2), C.c_int64(n), None, None, C.c_uint64(seed), C.c_uint64(off),
                             _p(mean), 2)
    idx = np.arange(n, dtype=np.uint64) + np.uint64(off)
    want2 = o.add_noise(clean, philox.snr_u01(seed, idx) * np.float32(70) + np.float32(50), philox.noise_eps(seed, idx, 11),
                        np.float64)
    assert rel_elem(sig2, want2) < SIG_TOL


def test_wide_posteriors_take_the_literal_fallback_of_the_kl(emu_elbo, qb):
    """Posteriors far outside the prior (means ~ N(0, 4), log-stds near their upper clip, draws scaled by 1.5): most KL
    samples leave the range of the five-moment estimator and go through the per-sample literal path, KL maps reach 4e5.
    There the reference's OWN float32 arithmetic is 5e-4 away from float64, so the float32 oracle is the yardstick: the
    kernel follows it to 1e-5; against float64 it is as far as the reference is."""
    ph = o.parse_params(_cfg())
    layer = qb.SignalGenerationLayer(_cfg(), True, True)
    params = _trainer(qb)._params_for(layer)
    n, S = 256, 70
    q, prior, sigma, data, mask = _elbo_batch(ph, n, 77)
    r = np.random.default_rng(1)
    q[:, 1], q[:, 3] = r.uniform(1.0, 4.0, n), r.uniform(1.0, 4.0, n)
    q[:, 0], q[:, 2] = r.normal(0, 4.0, n), r.normal(0, 4.0, n)
    eps = r.standard_normal((n, 2)).astype(np.float32)
    eps_kl = (r.standard_normal((n, S, 2)) * 1.5).astype(np.float32)
    got = _elbo(emu_elbo, params, q, sigma, data, mask, prior, eps, eps_kl, kl_samples=S)
    ref32 = o.elbo_and_grads(ph, q, sigma, data, mask, prior, eps, eps_kl, np.float32)
    ref64 = o.elbo_and_grads(ph, q, sigma, data, mask, prior, eps, eps_kl, np.float64)
    assert got['non_finite'] == 0 and np.isfinite(got['grad_q']).all() and float(np.abs(ref64['kl_map']).max()) > 1e5
    assert rel_elem(got['kl'], ref32['kl']) < 1e-5 and rel_max(got['kl_map'], ref32['kl_map']) < 1e-5
    assert rel_max(got['grad_q'], ref32['grad_q']) < 1e-5 and rel_elem(got['nll'], ref64['nll']) < 1e-5
    floor = max(rel_max(ref32['kl_map'], ref64['kl_map']), rel_max(ref32['grad_q'], ref64['grad_q']))
    assert floor > 1e-4                                     # the float32 reference itself misses the 1e-4 bar here ...
    assert rel_max(got['kl_map'], ref64['kl_map']) < 1.5 * floor and rel_max(got['grad_q'], ref64['grad_q']) < 1.5 * floor


def test_forward_kernel_at_and_beyond_the_edges_of_the_domain(emu, qb):
    """OEF = 0 (dw = 0: every Bessel argument 0), DBV = 0, the corners of the sampling box, OEF / DBV up to 1 and a negative
    DBV: value and gradient against the float64 oracle.  OEF = 2 (OEF x Hct > 0.40, not reachable with Hct 0.34): node 0
    of the float32 reference stops being exactly dead (domain note) -- the kernel follows the
    reference's float32 arithmetic there, not float64."""
    layer = qb.SignalGenerationLayer(_cfg(), True, True)
    ph = o.parse_params(_cfg())
    x = np.array([[0.0, 0.05], [1e-8, 0.05], [1e-4, 0.05], [0.84, 0.0], [0.04, 0.001], [0.84, 0.201], [1.0, 0.3], [0.5, 1.0],
                  [0.4, -0.05]], np.float32)
    g = np.random.default_rng(2).standard_normal((x.shape[0], 11)).astype(np.float32)
    sig, grad = _pair(emu, layer, x, g, grid=1, block=32)
    s64, g64 = o.forward_backward(ph, x, g, dtype=np.float64)
    assert np.isfinite(sig).all() and np.isfinite(grad).all()
    assert rel_elem(sig, s64) < 2e-6
    for i in range(x.shape[0]):
        assert rel_max(grad[i], g64[i]) < 2e-6, x[i]
    beyond = np.array([[2.0, 0.05], [1.5, 0.1]], np.float32)
    gb = np.ones((2, 11), np.float32)
    sig, grad = _pair(emu, layer, beyond, gb, grid=1, block=32)
    with np.errstate(all='ignore'):
        s32, g32 = o.forward_backward(ph, beyond, gb, dtype=np.float32)
        s64, _ = o.forward_backward(ph, beyond, gb, dtype=np.float64)
    assert rel_elem(sig, s32) < 2e-6 and rel_max(grad, g32) < 2e-6
    assert rel_elem(sig, s64) > 1e-3                        # ... where float64 is no longer the reference's oracle


@pytest.mark.parametrize('n', [1, 3, 33, 65])
def test_fused_elbo_ragged_batches(emu_elbo, qb, n):
    """A warp of k_elbo_pair takes units of 32 voxels, a lane pair two: batches that end inside a unit / inside a pair,
    with a masked voxel in the middle, against the float64 oracle (production and generic kernel)."""
    ph = o.parse_params(_cfg())
    layer = qb.SignalGenerationLayer(_cfg(), True, True)
    params = _trainer(qb)._params_for(layer)
    q, prior, sigma, data, mask = _elbo_batch(ph, n, 100 + n)
    mask[:] = 1
    if n > 2:
        mask[n // 2] = 0
    data = data * mask[:, None]
    r = np.random.default_rng(n)
    eps, eps_kl = r.standard_normal((n, 2)).astype(np.float32), r.standard_normal((n, 70, 2)).astype(np.float32)
    ref = o.elbo_and_grads(ph, q, sigma, data, mask, prior, eps, eps_kl, np.float64)
    for pair in (1, 0):
        got = _elbo(emu_elbo, params, q, sigma, data, mask, prior, eps, eps_kl, pair=pair, grid=1)
        assert rel_elem(got['nll'], ref['nll']) < GRAD_TOL and rel_elem(got['kl'], ref['kl']) < GRAD_TOL
        assert rel_max(got['grad_q'], ref['grad_q']) < GRAD_TOL and rel_max(got['grad_sigma'], ref['grad_sigma']) < GRAD_TOL
        assert not np.isnan(got['nll_map']).any() and not np.isnan(got['kl_map']).any()       # every voxel was written
        assert got['mask_sum'] == mask.sum()


def test_non_finite_inputs_stay_inside_their_voxel(emu, qb):
    """Two voxels share a warp in k_forward_pair (ballots over both halves pick the Bessel range): a NaN / Inf voxel must
    not change a single bit of its partner or of any other voxel (the reference relies on TerminateOnNaN)."""
    layer = qb.SignalGenerationLayer(_cfg(), True, True)
    x, g = _voxels(64, 5), np.ones((64, 11), np.float32)
    ref_s, ref_g = _pair(emu, layer, x, g, grid=1, block=32)
    xb = x.copy()
    xb[10], xb[21], xb[33] = [np.nan, 0.05], [0.4, np.inf], [np.inf, 0.05]
    s, gr = _pair(emu, layer, xb, g, grid=1, block=32)
    ok = [i for i in range(64) if i not in (10, 21, 33)]
    assert np.array_equal(s[ok], ref_s[ok]) and np.array_equal(gr[ok], ref_g[ok])
    assert np.isnan(s[10]).all() and np.isnan(s[21]).all() and np.isnan(gr[[10, 21, 33]]).all()


@pytest.mark.parametrize('multi', [False, True])
def test_fused_elbo_on_the_24_tau_grid(emu_elbo, qb, multi):
    """24 images, 16 distinct |tau| columns, spin echo at index 7: the generic k_elbo on the multi-group path (a lane per
    image, n_tau > 16) against the float64 oracle, single- and multi-image normalisation."""
    cfg = dict(_cfg(), tau_start='-0.028', tau_end='0.065', tau_step='0.004')
    ph = o.parse_params(cfg)
    layer = qb.SignalGenerationLayer(cfg, True, True)
    tr = qb.EncoderTrainer(cfg, student_t_df=200, multi_image_normalisation=multi, use_mvg=True, use_population_prior=False,
                           predict_log_data=False, seed=1)
    assert layer.n_tau == 24 and tr._se_idx == 7 and layer.params.n_cols == 16
    n, S = 48, 70
    r = np.random.default_rng(4)
    q = np.stack([r.normal(-0.3, 0.7, n), r.normal(0, 0.6, n), r.normal(-1.2, 0.7, n), r.normal(0, 0.6, n),
                  r.normal(0, 0.8, n)], -1).astype(np.float32)
    prior = (q + r.normal(0, 0.3, (n, 5))).astype(np.float32)
    sigma = np.exp(r.normal(np.log(0.05), 0.2, (n, 24))).astype(np.float32)
    truth = np.stack([r.uniform(0.1, 0.7, n), r.uniform(0.005, 0.15, n)], -1)
    data = (o.forward(ph, truth, dtype=np.float64) * 100 * (1 + 0.02 * r.standard_normal((n, 24)))).astype(np.float32)
    mask = (r.uniform(size=n) > 0.3).astype(np.float32)
    data *= mask[:, None]
    eps, eps_kl = r.standard_normal((n, 2)).astype(np.float32), r.standard_normal((n, S, 2)).astype(np.float32)
    ref = o.elbo_and_grads(ph, q, sigma, data, mask, prior, eps, eps_kl, np.float64, se_idx=7, multi_image_normalisation=multi)
    path = 0 if layer.params.sched_phases > 0 else 2
    got = _elbo(emu_elbo, tr._params_for(layer), q, sigma, data, mask, prior, eps, eps_kl, kl_samples=S, pair=0, path=path,
                grid=1)
    assert rel_elem(got['nll'], ref['nll']) < GRAD_TOL and rel_elem(got['kl'], ref['kl']) < GRAD_TOL
    assert rel_max(got['grad_q'], ref['grad_q']) < GRAD_TOL and rel_max(got['grad_sigma'], ref['grad_sigma']) < GRAD_TOL


def test_skinny_dense_kernels_of_the_heads(emu_enc):
    """k_dense_small_fwd_coop / k_dense_small_dgrad_coop (the 60 -> 5 and 60 -> 11 heads, model.py:176-223): coalesced
    row tiles through shared memory, shuffle-broadcast gradient rows, optional ReLU' mask on the result; ragged row
    counts; against float64 matrix products."""
    r = np.random.default_rng(14)
    for n, n_in, n_out, grid in ((1, 60, 5, 1), (95, 60, 11, 2), (300, 12, 16, 3), (257, 64, 1, 2)):
        x = r.standard_normal((n, n_in)).astype(np.float32)
        w = (r.standard_normal((n_out, n_in)) * 0.3).astype(np.float32)
        b = r.standard_normal(n_out).astype(np.float32)
        y = np.full((n, n_out), np.nan, np.float32)
        emu_enc.qb_emu_dense_small_forward(_p(x), _p(w), _p(b), n_in, n_out, C.c_int64(n), _p(y), grid)
        assert rel_max(y, x.astype(np.float64) @ w.astype(np.float64).T + b) < 1e-6
        g = r.standard_normal((n, n_out)).astype(np.float32)
        dx = np.full((n, n_in), np.nan, np.float32)
        emu_enc.qb_emu_dense_small_dgrad(_p(g), _p(w), None, n_in, n_out, C.c_int64(n), _p(dx), grid)
        want = g.astype(np.float64) @ w.astype(np.float64)
        assert rel_max(dx, want) < 1e-6
        act = r.standard_normal((n, n_in)).astype(np.float32)
        emu_enc.qb_emu_dense_small_dgrad(_p(g), _p(w), _p(act), n_in, n_out, C.c_int64(n), _p(dx), grid)
        assert rel_max(dx, want * (act > 0)) < 1e-6 and np.all(dx[act <= 0] == 0)


def test_signed_oef_option_restores_the_odd_continuation(tmp_path, qb):
    """-DQB_SIGNED_OEF (off in the library: the kernels it touches were not re-measured): with it the quadrature runs on
    |A| and the derivative sum takes the sign of A, so a NEGATIVE OEF -- which the reference's callers never produce --
    gives what TensorFlow's even 1 - J0 / odd J1 give; positive inputs keep their bits.  Without it the gradient of such a
    voxel is wrong (the known limitation stated in include/qbold.h and DESIGN.md section 4)."""
    gxx = shutil.which('g++')
    if not gxx:
        pytest.skip('g++ not available')
    libs = {}
    for tag, extra in (('plain', []), ('signed', ['-DQB_SIGNED_OEF'])):
        out = str(tmp_path / ('libforward_%s.so' % tag))
        subprocess.run([gxx, '-O1', '-std=c++17', '-pthread', '-DQB_HOST_EMU', '-fvisibility=hidden', '-fno-gnu-unique'] + extra +
                       ['-I', os.path.join(ROOT, 'tests', 'host_emu', 'shim'), '-I', os.path.join(ROOT, 'qbold_vi_b200', 'csrc'),
                        '-shared', '-fPIC', os.path.join(ROOT, 'tests', 'host_emu', 'forward_host.cpp'), '-o', out], check=True,
                       capture_output=True, timeout=900)
        libs[tag] = C.CDLL(out)
    layer = qb.SignalGenerationLayer(_cfg(), True, True)
    plain = type(layer.params)()
    C.memmove(C.byref(plain), C.byref(layer.params), C.sizeof(plain))
    plain.sched_phases = 0                                                           # -> column-group path
    ph = o.parse_params(_cfg())
    r = np.random.default_rng(3)
    neg = np.stack([-r.uniform(0.01, 0.84, 40), r.uniform(0.001, 0.2, 40)], -1).astype(np.float32)
    pos = _voxels(41, 9)
    g = r.standard_normal((41, 11)).astype(np.float32)
    s64, g64 = o.forward_backward(ph, neg, g[:40], dtype=np.float64)
    for run in (lambda lib, x, gg: _pair(lib, layer, x, gg), lambda lib, x, gg: _generic(lib, layer.params, x, gg, 11, path=0),
                lambda lib, x, gg: _generic(lib, plain, x, gg, 11, path=1)):
        sig, grad = run(libs['signed'], neg, g[:40])
        assert rel_elem(sig, s64) < SIG_TOL and rel_max(grad, g64) < GRAD_TOL
        a, b = run(libs['signed'], pos, g), run(libs['plain'], pos, g)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])             # positive OEF: not a bit changes
    _, grad_plain = _pair(libs['plain'], layer, neg, g[:40])
    assert not rel_max(grad_plain, g64) < 1e-2                                       # the limitation, as documented: wrong or non-finite
