"""CPU suite, part 3: the multi-rank host logic on gloo, world_size 2.

The fused CUDA loss cannot run here, so DataParallelTrainer gets a stand-in loss with the same contract
(normalised by the GLOBAL mask count); what is under test is sharding, the global mask sum, the single
flat-bucket all-reduce and replica consistency."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import qbold_oracle as o


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _standin_loss(q, sigma, data, mask, prior, mask_sum, seed=None):
    m = mask.reshape(-1)
    per_voxel = ((q.reshape(-1, 5) - prior.reshape(-1, 5)) ** 2).sum(-1) + \
        ((torch.log(sigma.reshape(m.shape[0], -1)) + 2.0) ** 2).sum(-1)
    loss = (per_voxel * m).sum() / mask_sum
    return loss, {'nll': loss.detach(), 'kl': loss.detach() * 0}


def _standin_tv(q, prior, mask, mask_sum):
    p = torch.sigmoid(torch.stack([q[..., 0], q[..., 2]], -1))
    both = (mask[:, :-1] > 0) & (mask[:, 1:] > 0)
    return ((p[:, :-1] - p[:, 1:]).abs() * both).sum() / mask_sum


def _make(seed=3, **kw):
    import qbold_vi_b200 as qb
    from qbold_vi_b200.encoder import Encoder
    from qbold_vi_b200.distributed import DataParallelTrainer
    cfg = o.default_config()
    torch.manual_seed(seed)
    enc = Encoder(no_units=12, no_intermediate_layers=2)
    tr = qb.EncoderTrainer(cfg, student_t_df=200, multi_image_normalisation=False, use_mvg=True,
                           use_population_prior=False, predict_log_data=False, seed=1)
    return enc, DataParallelTrainer(enc, tr, None, loss_fn=_standin_loss, tv_fn=_standin_tv, **kw)


def _batch():
    g = torch.Generator().manual_seed(11)
    data = torch.rand((4, 6, 5, 2, 11), generator=g) + 0.5
    mask = (torch.rand((4, 6, 5, 2, 1), generator=g) > 0.3).float()
    prior = torch.randn((4, 6, 5, 2, 5), generator=g) * 0.3
    return data * mask, mask, prior


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from qbold_vi_b200 import distributed as D
    r, w, dev = D.init_distributed('gloo')
    assert (r, w) == (rank, world) and dev.type == 'cpu'
    enc, dp = _make()
    data, mask, prior = _batch()
    lo, hi = D.shard_range(data.shape[0], rank, world)
    stats = dp.step(data[lo:hi], mask[lo:hi], prior[lo:hi])
    out.put((rank, dp.bucket.flat.clone().numpy(), stats.as_dict(),
             torch.cat([p.detach().reshape(-1) for p in enc.parameters()]).numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_is_a_balanced_partition():
    from qbold_vi_b200.distributed import shard_range
    for n in (0, 1, 7, 16, 1000003):
        for w in (1, 2, 3, 8):
            parts = [shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_step_equals_single_rank_step():
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([out.get(timeout=240) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference on the full batch
    enc, dp = _make()
    data, mask, prior = _batch()
    stats = dp.step(data, mask, prior)
    g_ref = dp.bucket.flat.numpy()
    w_ref = torch.cat([p.detach().reshape(-1) for p in enc.parameters()]).numpy()
    assert g_ref.shape[0] == sum(p.numel() for p in enc.parameters())
    for rank, g, st, w in res:
        assert np.max(np.abs(g - g_ref)) <= 1e-5 * np.max(np.abs(g_ref))       # all-reduced grads == full-batch grads
        assert np.max(np.abs(w - w_ref)) <= 1e-6                                # replicas apply the same update
        assert abs(st['loss'] - stats['loss']) <= 1e-5 * abs(stats['loss'])
        assert st['mask_sum'] == stats['mask_sum'] == float(mask.sum())
    assert np.array_equal(res[0][3], res[1][3])                                 # replicas stay bit-identical


def _graph_mode_worker(rank, world, port, out, tmp):
    """Host logic of the captured step (DataParallelTrainer(cuda_graph=...)) with the capture itself left out
    (graph_warmup=None): Philox key, schedule position and learning rate live in tensors and advance inside the step
    body; 'split' issues the collectives around the two halves."""
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from qbold_vi_b200 import distributed as D
    D.init_distributed('gloo')
    data, mask, prior = _batch()
    lo, hi = D.shard_range(data.shape[0], rank, world)
    d_, m_, p_ = data[lo:hi].contiguous(), mask[lo:hi].contiguous(), prior[lo:hi].contiguous()
    rows = []
    for mode in (True, 'split'):
        enc_e, dp_e = _make()
        enc_g, dp_g = _make(cuda_graph=mode, graph_warmup=None)
        worst = 0.0
        for i in range(4):
            a, b = dp_e.step(d_, m_, p_), dp_g.step(d_, m_, p_)
            for k in ('loss', 'nll', 'kl', 'smoothness', 'mask_sum'):
                worst = max(worst, abs(a[k] - b[k]) / max(abs(a[k]), 1e-6))
            assert abs(a['lr'] - b['lr']) < 1e-15
        # the device-resident scalars followed the host mirrors
        tr = dp_g.trainer
        want_seed = D._as_i64((tr._seed + D._GOLDEN * tr._calls) & 0xFFFFFFFFFFFFFFFF)
        seed_ok = int(dp_g._g['seed']) == want_seed and tr._calls == 4 and dp_g.step_no == 4
        sched_ok = abs(float(dp_g._g['sched'][0]) - dp_g.lr(3)) < 1e-9 and \
            abs(float(dp_g._g['sched'][1]) - (1.0 - dp_g.wd(3))) < 1e-7 and float(dp_g._g['t'][0]) == 4.0
        # checkpoint round trip into a fresh trainer of the same mode, then one more step on both
        path = os.path.join(tmp, 'g_%s_%d.pt' % (mode, rank))
        dp_g.save(path)
        enc_r, dp_r = _make(cuda_graph=mode, graph_warmup=None)
        dp_r.load(path)
        a, b = dp_e.step(d_, m_, p_), dp_r.step(d_, m_, p_)
        resumed = max(abs(a[k] - b[k]) / max(abs(a[k]), 1e-6) for k in ('loss', 'nll', 'smoothness'))
        w_e = torch.cat([p.detach().reshape(-1) for p in enc_e.parameters()])
        w_r = torch.cat([p.detach().reshape(-1) for p in enc_r.parameters()])
        rows.append((str(mode), worst, bool(seed_ok), bool(sched_ok), resumed, float((w_e - w_r).abs().max()),
                     w_r.numpy()))
    out.put((rank, rows))
    dist.barrier()
    dist.destroy_process_group()


def test_captured_step_host_logic_on_two_ranks(tmp_path):
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_graph_mode_worker, args=(r, 2, port, out, str(tmp_path))) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([out.get(timeout=300) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, rows in res:
        for mode, worst, seed_ok, sched_ok, resumed, w_err, _ in rows:
            assert worst < 1e-5 and resumed < 1e-5 and w_err < 1e-5, (rank, mode, worst, resumed, w_err)
            assert seed_ok and sched_ok, (rank, mode)
    for i in range(2):                                                          # replicas stay bit-identical
        assert np.array_equal(res[0][1][i][6], res[1][1][i][6])


def test_checkpoints_move_between_the_eager_and_the_captured_trainer(tmp_path):
    """A checkpoint written by the eager trainer resumes in the captured-step trainer and the other way round: the
    learning rate is re-aliased to the device schedule, counters and moments carry over (capture left out on CPU)."""
    data, mask, prior = _batch()
    enc_a, dp_a = _make()
    for _ in range(3):
        dp_a.step(data, mask, prior)
    path = str(tmp_path / 'eager.pt')
    dp_a.save(path)
    enc_b, dp_b = _make(seed=5, cuda_graph='split', graph_warmup=None)
    dp_b.load(path)
    assert dp_b.step_no == 3 and dp_b.opt.param_groups[0]['lr'].data_ptr() == dp_b._g['sched'].data_ptr()
    for _ in range(2):
        a, b = dp_a.step(data, mask, prior), dp_b.step(data, mask, prior)
        assert abs(a['loss'] - b['loss']) <= 1e-5 * abs(a['loss']) and abs(a['lr'] - b['lr']) < 1e-15
    path2 = str(tmp_path / 'captured.pt')
    dp_b.save(path2)
    enc_c, dp_c = _make(seed=7)
    dp_c.load(path2)
    a, c = dp_a.step(data, mask, prior), dp_c.step(data, mask, prior)
    assert dp_c.step_no == 6 and abs(a['loss'] - c['loss']) <= 1e-5 * abs(a['loss'])
    w_a = torch.cat([p.detach().reshape(-1) for p in enc_a.parameters()])
    w_c = torch.cat([p.detach().reshape(-1) for p in enc_c.parameters()])
    assert float((w_a - w_c).abs().max()) < 1e-5


class _SkipBody(Exception):
    pass


def _fake_cuda_graphs(monkeypatch, bodies):
    """torch.cuda.CUDAGraph / torch.cuda.graph stand-ins with the one property of a capture that matters to the host
    logic: the body of ``with torch.cuda.graph(g):`` is NOT executed (a trace function raises at its first line and
    __exit__ swallows it); ``g.replay()`` executes the body registered for the i-th graph created."""
    import sys
    created = []

    class FakeGraph:
        def __init__(self):
            self.index = len(created)
            created.append(self)

        def pool(self):
            return ('pool', self.index)

        def replay(self):
            bodies[self.index % len(bodies)]()

    class fake_graph:
        def __init__(self, graph, pool=None, **kw):
            self.graph, self.pool = graph, pool

        def __enter__(self):
            def tracer(frame, event, arg):
                raise _SkipBody()
            self._old = sys.gettrace()
            sys.settrace(lambda *a: None)                       # enable tracing, then trap the caller's next line
            sys._getframe(1).f_trace = tracer
            return self

        def __exit__(self, exc_type, exc, tb):
            sys.settrace(self._old)
            return exc_type is _SkipBody

    monkeypatch.setattr(torch.cuda, 'CUDAGraph', FakeGraph)
    monkeypatch.setattr(torch.cuda, 'graph', fake_graph)
    monkeypatch.setattr(torch.cuda, 'synchronize', lambda *a, **k: None)
    return created


@pytest.mark.parametrize('mode', [True, 'split'])
def test_capture_and_replay_control_flow_with_stand_in_graphs(monkeypatch, tmp_path, mode):
    """The capture / replay branches of DataParallelTrainer on CPU, with graph objects that behave like a capture on the
    host side (body skipped at capture, executed at replay): three eager warm-up steps, the capture call, replays, a new
    batch through the static buffers, a new batch SHAPE (new buffers, new capture), a checkpoint round trip (capture
    again) -- every step against the eager trainer."""
    data, mask, prior = _batch()
    enc_e, dp_e = _make()
    enc_g, dp_g = _make(cuda_graph=mode, graph_warmup=None)
    dp_g._g['warmup'] = 3                                       # capture after three eager steps, as on a GPU

    def front_full():
        dp_g._g['stats'] = dp_g._graph_body()

    def front_split():
        dp_g._g['stats'] = dp_g._split_front()

    bodies = [front_full] if mode is True else [front_split, dp_g._split_back]
    created = _fake_cuda_graphs(monkeypatch, bodies)

    def both(d, m, p, what):
        a, b = dp_e.step(d, m, p), dp_g.step(d, m, p)
        for k in ('loss', 'nll', 'smoothness', 'mask_sum'):
            assert abs(a[k] - b[k]) <= 1e-5 * max(abs(a[k]), 1e-6), (what, k, a[k], b[k])
        assert abs(a['lr'] - b['lr']) < 1e-15 and dp_g.step_no == dp_e.step_no

    for i in range(3):
        both(data, mask, prior, 'warm-up %d' % i)
    assert dp_g._g['graph'] is None and not created
    both(data, mask, prior, 'capture + first replay')
    assert dp_g._g['graph'] is not None and len(created) == (1 if mode is True else 2)
    if mode == 'split':
        assert dp_g._g['graph_back'] is created[1]
    for i in range(2):
        both(data, mask, prior, 'replay %d' % i)
    both(data * 1.02, mask, prior, 'new batch, same shape')
    assert len(created) == (1 if mode is True else 2)           # no new capture
    sd, sm, sp = dp_g.static_inputs()
    assert sd.data_ptr() != data.data_ptr() and torch.equal(sd, data * 1.02) and torch.equal(sm, mask)
    half = slice(0, 2)
    for i in range(5):                                          # new shape: three eager steps, capture, replay
        both(data[half].contiguous(), mask[half].contiguous(), prior[half].contiguous(), 'new shape %d' % i)
    assert len(created) == (2 if mode is True else 4)
    path = str(tmp_path / 'g.pt')
    dp_g.save(path)
    dp_g.load(path)
    assert dp_g._g['graph'] is None                             # moments are new tensors: capture again
    for i in range(5):
        both(data[half].contiguous(), mask[half].contiguous(), prior[half].contiguous(), 'after load %d' % i)
    assert len(created) == (3 if mode is True else 6)
    w_e = torch.cat([p.detach().reshape(-1) for p in enc_e.parameters()])
    w_g = torch.cat([p.detach().reshape(-1) for p in enc_g.parameters()])
    assert float((w_e - w_g).abs().max()) < 1e-5


def test_uint64_key_as_int64_bit_pattern():
    from qbold_vi_b200.distributed import _GOLDEN, _as_i64
    for u in (0, 1, (1 << 63) - 1, 1 << 63, (1 << 64) - 1, _GOLDEN):
        assert -(1 << 63) <= _as_i64(u) < (1 << 63) and _as_i64(u) % (1 << 64) == u
    t = torch.tensor([_as_i64((1 << 64) - 5)], dtype=torch.int64)
    t.add_(torch.tensor([_as_i64(_GOLDEN)], dtype=torch.int64))                  # wraps like the uint64 the kernel reads
    assert int(t) % (1 << 64) == ((1 << 64) - 5 + _GOLDEN) % (1 << 64)


def test_ranks_get_disjoint_host_cores():
    """pin_cores: every local rank takes its own slice of the cores this process may run on (QBOLD_PIN_CORES=0: off)."""
    import subprocess
    import sys
    if not hasattr(os, 'sched_getaffinity') or len(os.sched_getaffinity(0)) < 2:
        pytest.skip('needs sched_setaffinity and two cores')
    prog = ("import os, sys, json; sys.path.insert(0, %r)\n"
            "from qbold_vi_b200.distributed import pin_cores\n"
            "before = sorted(os.sched_getaffinity(0)); mine = pin_cores(int(sys.argv[1]), 2)\n"
            "print(json.dumps([before, mine, sorted(os.sched_getaffinity(0))]))\n"
            % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import json
    got = []
    for rank in (0, 1):
        out = subprocess.run([sys.executable, '-c', prog, str(rank)], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-500:]
        got.append(json.loads(out.stdout.strip().splitlines()[-1]))
    (b0, m0, a0), (b1, m1, a1) = got
    assert m0 == a0 and m1 == a1 and not set(a0) & set(a1) and set(a0) | set(a1) <= set(b0)
    assert len(a0) == len(a1) == len(b0) // 2
    env = dict(os.environ, QBOLD_PIN_CORES='0')
    out = subprocess.run([sys.executable, '-c', prog, '0'], capture_output=True, text=True, timeout=300, env=env)
    before, mine, after = json.loads(out.stdout.strip().splitlines()[-1])
    assert mine is None and before == after


def test_lazy_stats_convert_on_access_only():
    from qbold_vi_b200.distributed import LazyStats
    vals = torch.tensor([1.5, 2.5], dtype=torch.float64)
    st = LazyStats(['loss', 'kl'], vals, lr=0.1)
    assert st._host is None and st['lr'] == 0.1 and st._host is None              # host items never touch the tensor
    assert 'loss' in st and 'lr' in st and 'nll' not in st and st.get('nll', 7) == 7
    assert st['kl'] == 2.5 and st._host == [1.5, 2.5]
    assert st.as_dict() == {'lr': 0.1, 'loss': 1.5, 'kl': 2.5}


def test_encoder_parameter_count_matches_the_reference():
    from qbold_vi_b200.encoder import Encoder
    enc = Encoder()                                   # optimal.yaml: 60 units, 2 blocks, channel-wise gating
    assert sum(p.numel() for p in enc.parameters()) == 146176
    x = torch.rand(2, 5, 4, 3, 11) + 0.5
    q, q2, sg = enc(x)
    assert tuple(q.shape) == (2, 5, 4, 3, 5) and tuple(sg.shape) == (2, 5, 4, 3, 11) and bool((sg > 0).all())
    assert abs(float(sg.mean()) - 0.05) < 0.02                                   # bias = log(im_loss_sigma)
    # the 3x3x1 convs never mix z slices: a z-slab can be processed on its own (no halo when sharding by z)
    assert torch.allclose(enc(x[:, :, :, 1:2])[1], q2[:, :, :, 1:2], atol=1e-6)


def test_linear_schedule_matches_train_py():
    from qbold_vi_b200.distributed import LinearSchedule
    s = LinearSchedule(5e-3)
    assert s(0) == 5e-3 and abs(s(4000) - 5e-5) < 1e-12 and abs(s(2000) - (5e-3 + 5e-5) / 2) < 1e-12


def test_checkpoint_resume_continues_the_same_trajectory(tmp_path):
    """save() / load(): encoder weights, Adam moments, the LR / weight-decay schedule position and the gradient bucket
    aliasing survive a restart -- the resumed run takes the step the uninterrupted run takes."""
    data, mask, prior = _batch()
    enc_a, dp_a = _make()
    for _ in range(3):
        dp_a.step(data, mask, prior)
    path = str(tmp_path / 'final_model.pt')
    dp_a.save(path)
    ref = dp_a.step(data, mask, prior).as_dict()
    w_ref = torch.cat([p.detach().reshape(-1) for p in enc_a.parameters()])
    enc_b, dp_b = _make()
    dp_b.load(path)
    assert dp_b.step_no == 3
    assert all(p.grad.data_ptr() >= dp_b.bucket.flat.data_ptr() for p in dp_b.bucket.params)     # grads alias the bucket
    got = dp_b.step(data, mask, prior).as_dict()
    w_got = torch.cat([p.detach().reshape(-1) for p in enc_b.parameters()])
    assert abs(got['loss'] - ref['loss']) <= 1e-6 * abs(ref['loss']) and got['lr'] == ref['lr']
    assert torch.allclose(w_got, w_ref, rtol=0, atol=1e-7)
