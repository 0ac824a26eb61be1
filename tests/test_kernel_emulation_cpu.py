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
    subprocess.run([gxx, '-O1', '-std=c++17', '-pthread', '-DQB_HOST_EMU', '-I', os.path.join(ROOT, 'tests', 'host_emu', 'shim'),
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
