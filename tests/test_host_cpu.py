"""CPU suite, part 2: host logic and the C-ABI boundary (no compute calls: there is no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from oracle import qbold_oracle as o
from oracle import philox


@pytest.fixture(scope='module')
def qb():
    import qbold_vi_b200 as qb
    if not os.path.exists(qb._lib.LIB_PATH):
        qb.build_library()
    return qb


def test_library_exports_every_symbol_declared_in_the_header(qb):
    hdr = open(os.path.join(ROOT, 'include', 'qbold.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = set(re.findall(r'\b(qbold_[a-z0-9_]+)\s*\(', hdr))
    assert len(declared) >= 15
    handle = C.CDLL(qb._lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(handle, name), 'libqbold.so does not export %s' % name
    assert declared == set(qb._lib.EXPORTED_SYMBOLS), declared ^ set(qb._lib.EXPORTED_SYMBOLS)
    lib = qb._lib.lib()                       # also checks ABI version and sizeof(QboldParams)
    assert lib.qbold_abi_version() == 2


def test_params_block_matches_the_oracle(qb):
    cfg = o.default_config()
    ph = o.parse_params(cfg)
    layer = qb.SignalGenerationLayer(cfg, True, True)
    P = layer.params
    assert P.n_tau == 11 and P.n_cols == 8 and P.col_of_tau[2] == -1
    assert list(P.col_of_tau)[:5] == [0, 1, -1, 1, 0]                       # tau and -tau share a column
    assert np.array_equal(np.array(P.tau[:11], dtype=np.float32), ph.taus)
    assert np.float32(P.dw_k) == np.float32(o.dw_const(ph))
    assert np.float32(P.e_tissue) == o._exp(-ph.te * ph.r2t, np.float32)
    # 1 - (2 - e1) * e2 cancels ~2 bits: a 1-ulp difference between exp implementations shows up as ~1e-7 here
    assert abs(P.kappa - float(o.m_bld(ph, np.float64) * 0.775)) < 5e-8
    assert abs(P.kappa - float(o.m_bld(ph, np.float32) * np.float32(0.775))) < 3e-7
    B, _ = o.blood_b_of_tau(ph, np.float32)
    assert np.max(np.abs(np.array(P.blood_b[:11]) - B)) < 2e-6
    u = o.quad_nodes()
    assert np.array_equal(np.array(P.qu[:129], dtype=np.float32), u)
    W, _ = o.simpson_weights(u, np.float64)
    g = (2 + u.astype(np.float64)) * np.sqrt(1 - u.astype(np.float64)) / (3 * u.astype(np.float64) ** 2)
    c = W * g
    assert P.qc[0] == 0.0 and P.qc[128] == 0.0                              # node 0 dead in the value
    assert abs(P.node0_c - c[0]) / c[0] < 1e-6
    assert np.max(np.abs(np.array(P.qc[1:128]) - c[1:128]) / c[1:128]) < 1e-6
    assert np.max(np.abs(np.array(P.qd[0:128]) - (c * u)[0:128]) / (c * u)[0:128]) < 1e-6   # node 0 live in the derivative


def test_tau_24_grid_and_noise_table(qb):
    cfg = o.default_config()
    cfg.update(tau_start='-0.028', tau_end='0.065', tau_step='0.004')
    layer = qb.SignalGenerationLayer(cfg, True, True)
    assert layer.n_tau == 24 and layer.params.n_cols == 16
    ns = 1.0 - (np.abs(np.arange(-0.028, 0.065, 0.004)) * 3.0)
    assert np.max(np.abs(np.array(layer.params.norm_snr[:24]) - ns)) < 1e-6
    cfg.update(tau_start='0.0', tau_end='0.05', tau_step='0.01')            # 5 taus: no norm_snr (signals.py:117-121)
    layer = qb.SignalGenerationLayer(cfg, True, True)
    assert layer.n_tau == 5 and layer.params.norm_snr[0] == 0.0


def test_layer_argument_handling(qb):
    cfg = o.default_config()
    assert qb.SignalGenerationLayer(cfg, 'True', 'False')._include_blood is False
    with pytest.raises(ValueError):
        qb.SignalGenerationLayer(cfg, 'yes', True)
    both = qb.SignalGenerationLayer(cfg, True, True, misaligned_prob=0.1, variable_hct=True)   # signals.py:64-96
    assert both._variable_hct and both._misaligned_prob == 0.1
    import torch as _torch
    with pytest.raises(AssertionError):
        both(_torch.zeros(4, 2))                                            # needs (OEF, DBV, Hct) rows
    import torch
    layer = qb.SignalGenerationLayer(cfg, True, True)
    with pytest.raises(AssertionError):
        layer(torch.zeros(4, 3))
    with pytest.raises(qb.QboldError):                                      # no CPU path, ever
        qb.SignalGenerationLayer(dict(cfg, simulate_noise='False'), True, True)(torch.zeros(4, 2))
    assert abs(qb.SignalGenerationLayer.calculate_dw_static(1.0, 0.34, 2.67513e8, 3.0, 2.64e-7) - 301.743275) < 1e-5


def test_missing_library_fails_loudly(qb, monkeypatch, tmp_path):
    monkeypatch.setattr(qb._lib, '_lib', None)
    monkeypatch.setattr(qb._lib, 'LIB_PATH', str(tmp_path / 'libqbold.so'))
    with pytest.raises(qb.QboldError, match='no CPU or PyTorch fallback'):
        qb._lib.lib()


def test_config_two_tier_loading(qb):
    p = qb.load_system_parameters()
    assert isinstance(p['gamma'], str) and float(p['gamma']) == 2.67513e8 and p['simulate_noise'] == 'True'
    for k, v in o.default_config().items():
        assert float(p[k]) == float(v) if k not in ('simulate_noise', 'tau_weighted') else p[k] == v
    p['simulate_noise'] = 'False'                                            # callers mutate in place (train.py:256)
    assert p['simulate_noise'] == 'False'
    a = qb.optimal_arguments()
    assert (a.no_units, a.student_t_df, a.use_mvg, a.multi_image_normalisation, a.predict_log_data) == \
        (60, 200, True, False, False)
    assert a.save_directory == 'optimal' and a.name == 'optimal'            # extra yaml keys are added
    # typing rule of train.py:473-480: truthy defaults are cast, falsy defaults take the yaml value untyped
    args = qb.apply_yaml_overrides({'a': 1, 'b': 0.0, 'c': True}, {'a': '7', 'b': '3', 'c': 0, 'new': [1]})
    assert args == {'a': 7, 'b': '3', 'c': False, 'new': [1]}
    t = qb.EncoderTrainer(p, student_t_df=200)
    assert t._se_idx == 2


def test_philox_known_answers_and_streams():
    # Random123 kat_vectors, philox4x32-10
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = philox.philox4x32_10(*ctr, *key)
        assert tuple(int(x) for x in got) == want
    idx = np.arange(100000)
    n0, n1 = philox.normal_pair(11, idx, philox.STREAM_REPARAM)
    assert abs(n0.mean()) < 0.01 and abs(n0.std() - 1) < 0.01 and abs(np.corrcoef(n0, n1)[0, 1]) < 0.01
    u = philox.snr_u01(11, idx)
    assert 0 < u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 0.01
    # sharding invariance: the draw of voxel i does not depend on which slice it is generated in
    a = philox.reparam_eps(5, np.arange(1000, 2000))
    b = philox.reparam_eps(5, np.arange(0, 4000))[1000:2000]
    assert np.array_equal(a, b)
    for n in (1, 2, 529, 1000003):
        m = min(n, 5000)
        x = philox.feistel_permute(np.arange(n)[:m] if n > m else np.arange(n), n, 99)
        assert x.min() >= 0 and x.max() < n and len(np.unique(x)) == len(x)


def test_device_rng_source_on_the_host_matches_the_oracle_bit_for_bit(tmp_path):
    """csrc/rng.cuh (Philox4x32-10 and the 24-bit uniform, __host__ __device__ in the source) compiled for the host:
    the Random123 known answers, and word-for-word equality with oracle/philox.py for the counter layout the kernels use
    (index_lo, index_hi, stream, 0 | seed_lo, seed_hi), incl. indices beyond 2^32 and the stream ids."""
    import shutil
    import subprocess
    gxx = shutil.which('g++')
    if not gxx:
        pytest.skip('g++ not available')
    so = str(tmp_path / 'librng_emu.so')
    subprocess.run([gxx, '-O2', '-std=c++17', '-I', os.path.join(ROOT, 'qbold_vi_b200', 'csrc'), '-shared', '-fPIC',
                    os.path.join(ROOT, 'tests', 'host_emu', 'rng_host.cpp'), '-o', so], check=True, capture_output=True,
                   timeout=300)
    lib = C.CDLL(so)
    lib.qb_emu_stream_id.restype = C.c_uint32
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        out = (C.c_uint32 * 4)()
        lib.qb_emu_philox_raw((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out)
        assert tuple(out) == want
    assert [lib.qb_emu_stream_id(i) for i in range(5)] == [philox.STREAM_REPARAM, philox.STREAM_KL, philox.STREAM_SNR,
                                                            philox.STREAM_NOISE, philox.STREAM_MISALIGN]
    rng = np.random.default_rng(3)
    index = np.concatenate([np.arange(4096, dtype=np.uint64), rng.integers(0, 1 << 40, 4096, dtype=np.uint64),
                            np.uint64(1 << 32) + np.arange(-2, 3).astype(np.uint64)])
    for seed, stream in ((0, philox.STREAM_REPARAM), (0x9E3779B97F4A7C15, philox.STREAM_KL + 17),
                         (0xFFFFFFFFFFFFFFFF, philox.STREAM_MISALIGN)):
        words = np.empty((index.size, 4), np.uint32)
        u = np.empty((index.size, 4), np.float32)
        lib.qb_emu_philox(index.ctypes.data_as(C.c_void_p), C.c_int(index.size), C.c_uint32(stream), C.c_uint64(seed),
                          words.ctypes.data_as(C.c_void_p), u.ctypes.data_as(C.c_void_p))
        lo, hi, k0, k1 = philox._words(index, seed)
        want = np.stack(philox.philox4x32_10(lo, hi, np.uint32(stream), np.uint32(0), k0, k1), -1)
        assert np.array_equal(words, want)
        assert np.array_equal(u, philox.u01(want)) and u.min() > 0.0 and u.max() < 1.0
    # the Feistel shuffle of the streaming generator: a bijection of [0, n), identical to the oracle's
    for n, seed in ((1, 5), (2, 5), (529, 99), (6250000, 1), (1000003, 0xFEDCBA9876543210), ((1 << 33) + 7, 3)):
        m = int(min(n, 4000))
        i = np.arange(m, dtype=np.uint64) if n <= 4000 else rng.integers(0, n, m, dtype=np.uint64)
        out = np.empty(m, np.uint64)
        lib.qb_emu_feistel(i.ctypes.data_as(C.c_void_p), C.c_int(m), C.c_uint64(n), C.c_uint64(seed),
                           out.ctypes.data_as(C.c_void_p))
        assert np.array_equal(out, philox.feistel_permute(i, n, seed).astype(np.uint64)) and int(out.max()) < n
        if n <= 4000:
            assert len(np.unique(out)) == n


def test_static_lane_schedule_covers_every_node_once(qb):
    """QboldParams::sched_*: every (column, node>=1) pair is dealt to exactly one (pass, lane) with its Simpson
    weight, a lane keeps one column per phase, and the first-visit flags are consistent."""
    cfg = o.default_config()
    P = qb.SignalGenerationLayer(cfg, True, True).params
    nph, PL = P.sched_phases, 4
    assert 8 <= nph <= 10 and abs(P.tau_ref - 0.064) < 1e-7
    m = np.array(P.sched_m[:nph * PL * 32]).reshape(nph, PL, 32)
    w = np.array(P.sched_w[:nph * PL * 32]).reshape(nph, PL, 32)
    col = np.array(P.sched_col[:nph * 32]).reshape(nph, 32)
    u, c = np.array(P.qu[:129]), np.array(P.qc[:129])
    r = np.array(P.abs_tau[:8], dtype=np.float64) / P.tau_ref
    seen_nodes = {j: [] for j in range(8)}
    seen_slot = set()
    for ph in range(nph):
        for lane in range(32):
            j, first = int(col[ph, lane]) & 7, bool(col[ph, lane] & 0x80)
            assert first == ((lane, j) not in seen_slot)
            seen_slot.add((lane, j))
            for p in range(PL):
                if w[ph, p, lane] == 0.0:
                    continue
                k = int(round(m[ph, p, lane] / r[j] * 128 - 1e-5 * 128))
                assert abs(m[ph, p, lane] - r[j] * u[k]) < 1e-7 and w[ph, p, lane] == np.float32(c[k])
                seen_nodes[j].append(k)
        live = w[ph] != 0
        if live.any():
            assert P.sched_ph_min[ph] <= m[ph][live].min() and P.sched_ph_max[ph] >= m[ph][live].max()
    for j in range(8):
        assert sorted(seen_nodes[j]) == list(range(1, 128))          # node 0 handled in closed form, node 128 has weight 0
    # the point of the schedule: arguments inside one pass are close (ratio of max to min m, past the start-up phase)
    spread = [m[ph, p][w[ph, p] != 0].max() / m[ph, p][w[ph, p] != 0].min() for ph in range(1, nph) for p in range(PL)
              if (w[ph, p] != 0).any()]
    assert np.median(spread) < 1.4
    # more than 8 distinct |tau| (24-tau grid): no schedule, column-major path
    cfg.update(tau_start='-0.028', tau_end='0.065', tau_step='0.004')
    assert qb.SignalGenerationLayer(cfg, True, True).params.sched_phases == 0


def test_dataset_plumbing_shapes(qb):
    """prepare_dataset / prepare_synthetic_dataset semantics (train.py:17-104) on torch tensors."""
    import torch
    from qbold_vi_b200.data import FineTuneDataset, prepare_synthetic_dataset
    from qbold_vi_b200.encoder import Encoder
    torch.manual_seed(0)
    x, y = torch.rand(6000, 11), torch.rand(6010, 3)               # x shorter than y, as signals.py:283-287 can produce
    batches, (vx, vy) = prepare_synthetic_dataset(x, y, batch_size=4)
    assert tuple(vx.shape) == (1, 10, 10, 5, 11) and tuple(vy.shape) == (1, 10, 10, 5, 3)
    got = list(batches())
    assert sum(b[0].shape[0] for b in got) == 11 and tuple(got[0][0].shape) == (4, 10, 10, 5, 11)
    real = torch.rand(3, 70, 60, 8, 12) + 0.5
    real[..., -1] = (real[..., -1] > 1.0).float()
    ds = FineTuneDataset(real, Encoder(no_units=8, no_intermediate_layers=1), crop_size=25, training=True)
    (data, mask), tgt = next(iter(ds))
    assert tuple(data.shape) == (38, 25, 25, 8, 11) and tuple(mask.shape) == (38, 25, 25, 8, 1)      # 38*25*25*8 = 190 000 voxels
    assert tuple(tgt['predictions'].shape) == (38, 25, 25, 8, 6) and tuple(tgt['predicted_images'].shape) == (38, 25, 25, 8, 12)
    assert torch.equal(data, tgt['predicted_images'][..., :-1]) and bool((data[mask.expand_as(data) == 0] == 0).all())
    assert torch.equal(tgt['predictions'][..., -1:], mask)


def test_nifti_round_trip_and_subject_concatenation(tmp_path):
    """save_im_data (model.py:792-802): subjects concatenated along the last axis, NIfTI-1 single file, gzip."""
    import gzip
    import struct
    from qbold_vi_b200 import nifti
    rng = np.random.default_rng(0)
    maps = rng.standard_normal((3, 5, 4, 2, 1)).astype(np.float32)
    nifti.save_im_data(maps, str(tmp_path / 'oef'))
    arr, aff = nifti.load_nifti(str(tmp_path / 'oef.nii.gz'))
    assert arr.shape == (5, 4, 2, 3) and arr.dtype == np.float32
    for s in range(3):
        assert np.array_equal(arr[..., s], maps[s, ..., 0])
    assert np.array_equal(aff, np.eye(4, dtype=np.float32))
    raw = gzip.open(tmp_path / 'oef.nii.gz').read()
    assert len(raw) == 352 + 4 * arr.size and raw[344:348] == b'n+1\x00'
    assert struct.unpack_from('<8h', raw, 40)[:5] == (4, 5, 4, 2, 3) and struct.unpack_from('<h', raw, 70)[0] == 16
    assert np.frombuffer(raw, np.float32, 2, 352).tolist() == [arr[0, 0, 0, 0], arr[1, 0, 0, 0]]   # x fastest
    vol = rng.integers(0, 255, (4, 3, 2)).astype(np.uint8)
    nifti.save_nifti(vol, str(tmp_path / 'm.nii'), affine=np.diag([2.0, 2.0, 3.0, 1.0]))
    back, aff = nifti.load_nifti(str(tmp_path / 'm.nii'))
    assert np.array_equal(back, vol) and aff[2, 2] == 3.0


def test_encoder_host_paths_on_cpu(qb):
    """The encoder's routing helpers leave CPU tensors on the plain torch layers (the kernels are CUDA-only), and the
    stream-1-only forward used by pre-training equals output 0 of the full forward."""
    import torch
    from qbold_vi_b200.encoder import Encoder, dense, gate_mix
    torch.manual_seed(0)
    enc = Encoder(no_units=12, no_intermediate_layers=2)
    x = torch.rand(2, 5, 4, 3, 11) * 100.0 + 10.0
    full = enc(x)
    assert tuple(full[0].shape) == (2, 5, 4, 3, 5) and tuple(full[2].shape) == (2, 5, 4, 3, 11)
    assert torch.allclose(enc.forward_voxelwise(x), full[0], atol=1e-6)
    assert enc.supports_voxelwise_fused() and not Encoder(activation='gelu').supports_voxelwise_fused()
    assert not Encoder(no_units=80).supports_voxelwise_fused()
    h = torch.randn(7, 12, requires_grad=True)
    y = dense(enc.blocks[0].pointwise, h, relu=True)
    assert torch.allclose(y, torch.relu(enc.blocks[0].pointwise(h)))
    skip, r, z = torch.randn(3, 12), torch.randn(3, 12), torch.randn(3, 12)
    g = torch.sigmoid(z - 3.0)
    assert torch.allclose(gate_mix(skip, r, z, -3.0), skip * (1 - g) + r * g)
    # parameter count of the optimal.yaml encoder (what the gradient all-reduce carries)
    assert sum(p.numel() for p in Encoder(no_units=60, no_intermediate_layers=2).parameters()) == 146176


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours): one JSON line with the contract keys."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup',
                          '0'], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
                'vs_baseline', 'dtype', 'data', 'config', 'impl', 'cpu_baseline', 'e2e', 'gpu_launches'):
        assert key in line, key
    assert line['impl'] == 'reference' and line['unit'] == 'voxel-signals/s' and line['higher_is_better'] is True
    assert line['value'] > 0 and line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] >= 1
    assert line['e2e']['h2d_bytes_per_step'] == 0 and 'workload' in line['config']


def test_bench_fails_loudly_without_a_gpu():
    """The product arm has no CPU fallback: without CUDA it must stop with a clear message, not print a number."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--steps', '1'], capture_output=True, text=True,
                         timeout=300, cwd=root)
    assert out.returncode != 0 and 'no CPU fallback' in (out.stderr + out.stdout)
    assert '"value"' not in out.stdout


@pytest.mark.parametrize('case', ['ok', 'raises', 'hangs'])
def test_bench_prints_its_line_whatever_the_captured_step_trial_does(case):
    """bench.py runs the captured (CUDA graph) training step last and under a deadline: a trial that succeeds updates
    the training_step leg, one that raises or never returns leaves the eager numbers in the line; exit code 0 always."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    prog = (
        "import sys, time, types, atexit; sys.path.insert(0, %r); import bench\n"
        "atexit.register(lambda: sys.stderr.write('EXIT-HOOK-RAN\\n'))\n"
        "line = {'metric': 'm', 'training_step': {'ms_per_step': 4.0, 'ms_encoder_fwd_bwd': 3.0}}\n"
        "args = types.SimpleNamespace(graph_deadline=2)\n"
        "def trial(mode):\n"
        "    case = %r\n"
        "    if case == 'raises': raise RuntimeError('capture failed')\n"
        "    if case == 'hangs': time.sleep(60)\n"
        "    return {'ms_per_step': 3.5, 'voxel_signals_per_s': 1.0, 'ms_host_enqueue_per_step': 0.05, 'loss': -1.0}\n"
        "bench.finish_with_graph_trial(line, line['training_step'], trial, 0, 1, args)\n" % (root, case))
    out = subprocess.run([sys.executable, '-c', prog], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr[-800:]
    assert 'EXIT-HOOK-RAN' in out.stderr                  # exit hooks of the harness run on every way out
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith('{')]
    assert len(lines) == 1
    t = json.loads(lines[0])['training_step']
    if case == 'ok':
        assert t['ms_per_step'] == 3.5 and t['captured_step'] == {'mode': 'full', 'status': 'ok'}
        assert 'CUDA graph' in t['launch_mode']
    else:
        assert t['ms_per_step'] == 4.0 and t['captured_step']['mode'] == 'full'
        assert ('failed: RuntimeError' if case == 'raises' else 'deadline') in t['captured_step']['status']


@pytest.fixture(scope='module')
def bessel_emu(tmp_path_factory):
    """csrc/bessel.cuh compiled for the host (tests/host_emu/bessel_host.cpp, -DQB_HOST_EMU)."""
    import shutil
    import subprocess
    gxx = shutil.which('g++')
    if not gxx:
        pytest.skip('g++ not available')
    out = str(tmp_path_factory.mktemp('emu') / 'libbessel_emu.so')
    subprocess.run([gxx, '-O2', '-std=c++17', '-DQB_HOST_EMU', '-I', os.path.join(ROOT, 'qbold_vi_b200', 'csrc'), '-shared',
                    '-fPIC', os.path.join(ROOT, 'tests', 'host_emu', 'bessel_host.cpp'), '-o', out], check=True,
                   capture_output=True, timeout=300)
    return C.CDLL(out)


def _emu_bessel(lib, x, which):
    x = np.ascontiguousarray(x, np.float32)
    a, b = np.empty_like(x), np.empty_like(x)
    lib.qb_emu_bessel(x.ctypes.data_as(C.c_void_p), C.c_int(x.size), C.c_int(which), a.ctypes.data_as(C.c_void_p),
                      b.ctypes.data_as(C.c_void_p))
    return a, b


def test_device_bessel_source_on_the_host_against_scipy(bessel_emu):
    """The Bessel kernels of the hot path (csrc/bessel.cuh, the source the GPU kernels are built from) evaluated on the
    host: (1 - J0, J1) against scipy in each range INCLUDING the overlap a straddling warp pass uses (mid on [2, 9],
    big from 6.5), at the error bars the header states; the reference's own float32 Bessel (Cephes j0f, 1.9e-7) is the
    yardstick -- the signal tolerance of 1e-5 leaves a factor of 20."""
    import scipy.special as sp
    bars = {'production selection': (0.0, 40.0, 0, 6e-7, 4.5e-7), 'small': (0.0, 3.0, 1, 2.5e-7, 2.5e-7),
            'mid': (2.0, 9.0, 2, 5e-7, 4e-7), 'big': (6.5, 40.0, 3, 6e-7, 4.5e-7), 'big, far': (40.0, 400.0, 3, 1.2e-6, 1.5e-6)}
    for name, (lo, hi, which, bar0, bar1) in bars.items():
        x = np.linspace(lo, hi, 200001).astype(np.float32)
        omj0, j1 = _emu_bessel(bessel_emu, x, which)
        x64 = x.astype(np.float64)
        assert np.abs(omj0 - (1.0 - sp.j0(x64))).max() < bar0, name
        assert np.abs(j1 - sp.j1(x64)).max() < bar1, name
    # small arguments carry the largest quadrature weights (~ 1 / u^2): RELATIVE accuracy there, no cancellation
    x = np.linspace(1e-6, 0.05, 2001).astype(np.float32)
    omj0, j1 = _emu_bessel(bessel_emu, x, 0)
    z = x.astype(np.float64) ** 2
    series = z / 4 - z * z / 64 + z ** 3 / 2304
    assert np.max(np.abs(omj0 - series) / series) < 2.5e-7
    assert np.max(np.abs(j1 - sp.j1(x.astype(np.float64))) / sp.j1(x.astype(np.float64))) < 2.5e-7
    x = np.linspace(0.05, 3.0, 20001).astype(np.float32)
    omj0, _ = _emu_bessel(bessel_emu, x, 0)
    want = 1.0 - sp.j0(x.astype(np.float64))
    assert np.max(np.abs(omj0 - want) / want) < 4e-7
    assert _emu_bessel(bessel_emu, np.zeros(1, np.float32), 0)[0][0] == 0.0       # node 0 of a dead voxel stays exactly 0


def test_packed_quadrature_steps_equal_the_scalar_kernels(bessel_emu):
    """acc_small2 / acc_mid2 / acc_big2 (fma.rn.f32x2 on register pairs; emulated per half on the host) accumulate
    w (1 - J0) and (w m) J1 -- the same values the scalar kernels give, to the last float32 rounding of the products."""
    import scipy.special as sp
    w, A = np.float32(0.37), np.float32(13.0)
    for which, lo, hi in ((1, 0.0, 3.0), (2, 2.0, 9.0), (3, 6.5, 40.0)):
        x = np.linspace(lo, hi, 100000).astype(np.float32)
        acc_i, acc_b = np.empty_like(x), np.empty_like(x)
        bessel_emu.qb_emu_acc2(x.ctypes.data_as(C.c_void_p), C.c_int(x.size // 2), C.c_int(which), C.c_float(w), C.c_float(A),
                               acc_i.ctypes.data_as(C.c_void_p), acc_b.ctypes.data_as(C.c_void_p))
        omj0, j1 = _emu_bessel(bessel_emu, x, which)
        assert np.abs(acc_i - w * omj0).max() <= 1.2e-7                          # one rounding of the product
        x64 = x.astype(np.float64)
        m = (x / A).astype(np.float64)
        want_b = float(w) * x64 * sp.j1(x64) if which == 1 else float(w) * m * sp.j1(x64)
        assert np.abs(acc_b - want_b).max() < 6e-7
        assert np.abs(acc_i - float(w) * (1.0 - sp.j0(x64))).max() < 3e-7


def test_ctypes_signatures_match_the_header_prototypes(qb):
    """Every prototype of include/qbold.h against the ctypes signature the Python side binds it with: same number of
    parameters, pointer / int32 / int64 / uint64 / float / double in the same positions (a drift here is a silent
    ABI bug, not a link error)."""
    hdr = re.sub(r'/\*.*?\*/', '', open(os.path.join(ROOT, 'include', 'qbold.h')).read(), flags=re.S)
    protos = re.findall(r'\b([A-Za-z_][\w\s\*]*?)\b(qbold_[a-z0-9_]+)\s*\(([^)]*)\)\s*;', hdr)
    assert len(protos) == len(qb._lib._SIGNATURES) >= 50

    def kind(decl):
        d = decl.strip()
        if d in ('void', ''):
            return None
        if '*' in d:
            return 'ptr'
        for k, v in (('uint64_t', 'u64'), ('int64_t', 'i64'), ('int32_t', 'i32'), ('double', 'f64'), ('float', 'f32'),
                     ('int', 'i32')):
            if re.search(r'\b%s\b' % k, d):
                return v
        raise AssertionError('unhandled parameter type: %r' % d)

    def ckind(a):
        for t, v in ((C.c_uint64, 'u64'), (C.c_int64, 'i64'), (C.c_int32, 'i32'), (C.c_double, 'f64'), (C.c_float, 'f32'),
                     (C.c_void_p, 'ptr'), (C.c_char_p, 'ptr')):
            if a is t:
                return v
        assert hasattr(a, '_type_') and not isinstance(a._type_, str), a          # POINTER(struct) / POINTER(c_float)
        return 'ptr'

    for ret, name, args in protos:
        res, argtypes = qb._lib._SIGNATURES[name]
        want = [k for k in (kind(a) for a in args.split(',')) if k]
        assert [ckind(a) for a in argtypes] == want, name
        want_ret = 'ptr' if '*' in ret else kind(ret)
        assert (None if res is None else ckind(res)) == want_ret, name


def test_every_entry_point_rejects_garbage_before_touching_the_device(qb):
    """NULL pointers, negative sizes and zero scalars into every int-returning entry point: a negative QBOLD_E* code and
    a message that names the function -- decided on the host, so it holds (and is tested) without a GPU.  The launch
    counter does not move."""
    lib = qb._lib.lib()
    sizes = {'qbold_abi_version', 'qbold_params_sizeof', 'qbold_encoder_mlp_blob_floats', 'qbold_dense_tc_packed_floats'}
    before = qb.launch_count()
    checked = 0
    for name, (restype, argtypes) in qb._lib._SIGNATURES.items():
        if restype is not C.c_int or name in sizes:
            continue
        args = []
        for a in argtypes:
            if a in (C.c_void_p, C.c_char_p) or (hasattr(a, '_type_') and not isinstance(a._type_, str)):
                args.append(None)
            elif a in (C.c_float, C.c_double):
                args.append(0.0)
            else:
                args.append(-1)
        rc = getattr(lib, name)(*args)
        msg = lib.qbold_last_error().decode()
        assert rc in (-1, -3), (name, rc, msg)                                   # QBOLD_EINVAL / QBOLD_EUNSUPPORTED
        assert msg.startswith(name) or (name.endswith('_add') and msg.startswith(name[:-4])), (name, msg)
        checked += 1
    assert checked >= 40 and qb.launch_count() == before


def test_header_is_plain_c_and_the_c_demo_links(qb, tmp_path):
    """The boundary is a C ABI: include/qbold.h must parse as strict C99 and as C++, and examples/c_abi_demo.c must
    compile with a C compiler and link against libqbold.so (it is RUN by the GPU suite: test_c_abi_from_plain_c)."""
    import shutil
    import subprocess
    gcc, gxx = shutil.which('gcc'), shutil.which('g++')
    if not gcc or not gxx:
        pytest.skip('gcc / g++ not available')
    hdr = os.path.join(ROOT, 'include', 'qbold.h')
    subprocess.run([gcc, '-std=c99', '-pedantic', '-Wall', '-Werror', '-fsyntax-only', '-x', 'c', hdr], check=True,
                   capture_output=True, timeout=120)
    subprocess.run([gxx, '-std=c++11', '-Wall', '-Werror', '-fsyntax-only', '-x', 'c++', hdr], check=True,
                   capture_output=True, timeout=120)
    cuda = os.environ.get('CUDA_HOME', '/usr/local/cuda')
    if not os.path.exists(os.path.join(cuda, 'include', 'cuda_runtime.h')):
        pytest.skip('CUDA toolkit headers not available')
    qb._lib.lib()                                                    # the library exists (built in-tree)
    libdir = os.path.join(ROOT, 'qbold_vi_b200')
    res = subprocess.run([gcc, '-std=c99', '-Wall', '-I', os.path.join(ROOT, 'include'), '-I', os.path.join(cuda, 'include'),
                          os.path.join(ROOT, 'examples', 'c_abi_demo.c'), '-L', libdir, '-lqbold', '-L',
                          os.path.join(cuda, 'lib64'), '-lcudart', '-lm', '-o', str(tmp_path / 'c_abi_demo')],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-1500:]


def test_dlpack_unpacking_without_a_gpu(qb):
    """The ctypes DLPack consumer: capsule and __dlpack__ producers, shape / dtype / contiguity checks, single use,
    and the loud refusal of host memory (there is no CPU path)."""
    import torch
    from torch.utils import dlpack as tdl
    t = torch.arange(24, dtype=torch.float32).reshape(4, 3, 2)
    with qb.DLPackView(tdl.to_dlpack(t), allow_host=True) as v:
        assert v.shape == (4, 3, 2) and v.dtype == 'f4' and v.numel == 24 and v.ptr == t.data_ptr()
    with qb.DLPackView(t[1:], allow_host=True) as v:                         # __dlpack__ protocol, offset view
        assert v.shape == (3, 3, 2) and v.ptr == t[1:].data_ptr()
    cap = tdl.to_dlpack(t)
    qb.DLPackView(cap, allow_host=True).release()
    with pytest.raises(qb.dlpack.DLPackError, match='already consumed'):
        qb.DLPackView(cap, allow_host=True)
    with pytest.raises(qb.dlpack.DLPackError, match='C-contiguous'):
        qb.DLPackView(t.transpose(0, 1), allow_host=True)
    with pytest.raises(qb.dlpack.DLPackError, match='unsupported DLPack dtype'):
        qb.DLPackView(t.double(), allow_host=True)
    with pytest.raises(qb.dlpack.DLPackError, match='expected a f4'):
        qb.DLPackView(t.int(), allow_host=True)
    with pytest.raises(qb.dlpack.DLPackError, match='CUDA device memory only'):
        qb.DLPackView(t)
    with pytest.raises(qb.dlpack.DLPackError):
        qb.DLPackView(object())


def test_encoder_matches_the_reference_source_network(qb):
    """EncoderTrainer.create_encoder (model.py:122-223), executed from the reference's own source over the Keras shim
    (oracle/make_golden.py, fixture ref_shim_encoder.npz): same weights -> same three outputs and the same gradient for
    every kernel and bias.  Pins normalise_data, the shared pointwise layer, the gated residual blocks, both heads."""
    import torch
    from conftest import golden, load_reference_encoder_weights, reference_encoder_grad
    from qbold_vi_b200.encoder import Encoder
    fix = golden('ref_shim_encoder.npz')
    enc = Encoder(no_units=60, no_intermediate_layers=2, activation='relu', initial_im_sigma=0.05,
                  multi_image_normalisation=False, channelwise_gating=True, gate_offset=float(fix['gate_offset']),
                  resid_init_std=0.1, no_ip_images=11, se_idx=2, use_mvg=True)
    params = load_reference_encoder_weights(enc, fix)
    outs = enc(torch.as_tensor(fix['data']))
    for o, key in zip(outs, ('out_voxelwise', 'out_spatial', 'out_sigma')):
        ref = torch.as_tensor(fix[key])
        assert o.shape == ref.shape and float((o - ref).abs().max()) <= 2e-5 * float(ref.abs().max()), key
    sum((o * torch.as_tensor(fix['w_out%d' % i])).sum() for i, o in enumerate(outs)).backward()
    for i, (w, b, kind) in enumerate(params):
        gw, gb = reference_encoder_grad(fix, i, kind)
        assert float((w.grad - gw).abs().max()) <= 1e-4 * float(gw.abs().max()) + 1e-7, ('kernel', i)
        assert float((b.grad - gb).abs().max()) <= 1e-4 * float(gb.abs().max()) + 1e-7, ('bias', i)
