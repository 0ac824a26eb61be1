"""Voxel-sharded data parallelism for the hot path: one process per GPU.

* forward / generation / inference: contiguous voxel shards, no communication;
* training: every loss of the reference divides by the GLOBAL ``sum(mask)`` (model.py:566,663,753), so the mask
  count is all-reduced first (it is an input) and handed to the fused kernel as ``inv_mask_sum``; the only
  per-step collectives are ONE all-reduce of the flat encoder-gradient bucket (146 176 floats for optimal.yaml,
  latency-bound over NVLink 5 / NVSwitch) and one of the 4 loss partial sums.  Replicas stay bit-identical
  because they start from the same seed and apply the same reduced gradient.
"""
from __future__ import annotations

import os
import time

import torch

from ._lib import on_device as _on_device
import torch.distributed as dist


def init_distributed(backend=None):
    """torchrun-style rendezvous (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*); returns (rank, world, device)."""
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    use_cuda = torch.cuda.is_available()
    pin_cores(local, int(os.environ.get('LOCAL_WORLD_SIZE', world)))
    device = torch.device('cuda', local) if use_cuda else torch.device('cpu')
    if use_cuda:
        torch.cuda.set_device(device)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29512')
        backend = backend or ('nccl' if use_cuda else 'gloo')
        if backend == 'nccl':
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=device)
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, device


def pin_cores(local_rank, local_world):
    """Give every rank of a node its own slice of the host cores (QBOLD_PIN_CORES=0 disables).  The synchronous step
    waits for the slowest rank, so a rank whose launch thread gets descheduled behind another rank's helper threads
    costs all of them; disjoint core sets keep the arrival skew of the gradient all-reduce down."""
    if local_world <= 1 or os.environ.get('QBOLD_PIN_CORES', '1') != '1' or not hasattr(os, 'sched_setaffinity'):
        return None
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = len(cores) // local_world
        if per < 1:
            return None
        mine = cores[local_rank * per:(local_rank + 1) * per]
        os.sched_setaffinity(0, mine)
        return mine
    except OSError:
        return None


def world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def shard_range(n, rank, world):
    """Contiguous, balanced shard [lo, hi) of n units (voxels or volumes) for `rank`."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_reduce_sum_(t):
    if world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def global_mask_sum(mask):
    """sum(mask) over all ranks, as a Python float (one tiny all-reduce; the mask is an input, so this can
    be issued before the step).  Synchronises with the host: the training step uses global_mask_sum_device."""
    return float(global_mask_sum_device(mask).item())


def global_mask_sum_device(mask):
    """sum(mask) over all ranks as a 1-element float64 DEVICE tensor: the all-reduce is enqueued on the stream and
    the value is consumed by the kernels through a device pointer (qbold_elbo_fused_dev / qbold_smoothness_dev),
    so the step never waits for it on the host."""
    return all_reduce_sum_(mask.sum(dtype=torch.float64).reshape(1))


class LazyStats(dict):
    """Per-step statistics that stay on the device until somebody reads them: ``stats['loss']`` converts (and thereby
    synchronises) on access, so a training loop that logs every k-th step has no host sync on the others."""

    def __init__(self, names, values, **host_items):
        super().__init__(host_items)
        self._names, self._values, self._host = list(names), values, None

    def _fetch(self):
        if self._host is None:
            self._host = self._values.tolist()
        return self._host

    def __getitem__(self, key):
        if key in self._names:
            return self._fetch()[self._names.index(key)]
        return super().__getitem__(key)

    def get(self, key, default=None):
        return self[key] if (key in self._names or key in self.keys()) else default

    def __contains__(self, key):
        return key in self._names or super().__contains__(key)

    def as_dict(self):
        d = dict(self)
        d.update(zip(self._names, self._fetch()))
        return d


def _adam(params, **kw):
    """Adam with the single-kernel (fused) update on CUDA parameters; the multi-tensor default elsewhere (CPU tests)."""
    params = list(params)
    kw.setdefault('eps', 1e-7)                       # tf.keras / tfa Adam(W) epsilon (train.py:309-311), not torch's 1e-8
    if params and all(p.is_cuda for p in params):
        try:
            return torch.optim.Adam(params, fused=True, **kw)
        except (RuntimeError, TypeError):
            pass
    return torch.optim.Adam(params, **kw)


_GOLDEN = 0x9E3779B97F4A7C15                     # per-call seed stride of EncoderTrainer (model._next_seed)


def _as_i64(u):
    """The int64 with the bit pattern of the uint64 `u` (torch has no uint64 arithmetic; the kernel reads uint64)."""
    return u - (1 << 64) if u >= (1 << 63) else u


class FlatGradBucket:
    """One contiguous float32 buffer aliasing every parameter's .grad: a single all-reduce per step."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            # same (dense, possibly permuted) strides as the parameter, e.g. channels_last_3d conv weights
            p.grad = self.flat[off:off + p.numel()].as_strided(p.size(), p.stride())
            off += p.numel()

    def realias(self):
        """Point every parameter's .grad back into the flat buffer (after load_state_dict / anything that replaced it)."""
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].as_strided(p.size(), p.stride())
            off += p.numel()

    def zero_(self):
        self.flat.zero_()

    def all_reduce_(self):
        # gradients of a loss that is already normalised by the global mask count: SUM, no averaging
        return all_reduce_sum_(self.flat)


class LinearSchedule:
    """LRSchedule of train.py:287-306: initial value until step 0, then linear decay to initial/100 over
    40 x 100 steps (used for both the learning rate and the AdamW weight decay)."""

    def __init__(self, initial, steps_per_epoch=100, epochs=40.0):
        self.initial = float(initial)
        self.rate = (self.initial / 1e2 - self.initial) / (epochs * steps_per_epoch)

    def __call__(self, step):
        return self.initial + self.rate * step if step > 0 else self.initial


class DataParallelTrainer:
    """Fine-tuning step of the reference (train.py:285-376) over voxel/volume shards.

    loss = NLL + kl_weight * KL + smoothness_weight * TV, each normalised by the global mask count.
    ``loss_fn(q, sigma, data, mask, prior, mask_sum) -> (loss, info)`` defaults to the fused sm_100a kernel
    (EncoderTrainer.fused_elbo) and ``tv_fn(q, prior, mask, mask_sum) -> tv`` to the stencil kernel
    (EncoderTrainer.smoothness_loss); tests of the host logic inject stand-ins."""

    def __init__(self, encoder, trainer, signal_layer, ft_lr=5e-3, adamw_decay=2e-4, smoothness_weight=5.0,
                 kl_weight=1.0, kl_samples=70, loss_fn=None, tv_fn=None, cuda_graph=False, graph_warmup=3):
        self.encoder, self.trainer, self.layer = encoder, trainer, signal_layer
        self.smoothness_weight, self.kl_weight, self.kl_samples = smoothness_weight, kl_weight, kl_samples
        self.bucket = FlatGradBucket(encoder.parameters())
        self.lr, self.wd = LinearSchedule(ft_lr), LinearSchedule(adamw_decay)
        self.decay = adamw_decay > 0.0
        self.step_no = 0
        self._comm_stream = None
        self.loss_fn = loss_fn or self._fused_loss
        self.tv_fn = tv_fn or self._tv
        # cuda_graph=True: the whole step (encoder, fused ELBO, TV, backward, the three NCCL all-reduces, AdamW) is
        # captured once and replayed with ONE cudaGraphLaunch per step.  Everything that changes from step to step
        # lives on the device: the Philox key, the schedule position (learning rate, weight decay), Adam's counter.
        # cuda_graph='split': the same, with the NCCL calls kept OUT of the capture -- graph A (encoder, losses, backward),
        # eager all-reduce of the gradient bucket, graph B (weight decay, Adam, statistics); the mask count is reduced
        # before A and the statistics after B.  Two graph launches and three eager collectives per step.
        if cuda_graph not in (False, True, None, 'full', 'split'):
            raise ValueError("cuda_graph must be False, True ('full') or 'split'")
        self.cuda_graph = bool(cuda_graph)
        self.graph_mode = None if not cuda_graph else ('split' if cuda_graph == 'split' else 'full')
        self._g = None
        if self.cuda_graph:
            # (a custom loss_fn must take ``seed=`` -- the device Philox key -- and launch nothing that cannot be captured)
            dev = self.bucket.flat.device
            if dev.type != 'cuda' and graph_warmup is not None:
                raise ValueError('cuda_graph needs the encoder on a CUDA device (graph_warmup=None never captures: the '
                                 'step then runs eagerly on the device-resident scalars, which is what the CPU tests of '
                                 'the host logic use)')
            f32 = dict(dtype=torch.float32, device=dev)
            self._g = {'warmup': float('inf') if graph_warmup is None else int(graph_warmup), 'calls': 0, 'graph': None,
                       'shapes': None,
                       'seed': torch.zeros(1, dtype=torch.int64, device=dev),
                       'golden': torch.tensor([_as_i64(_GOLDEN)], dtype=torch.int64, device=dev),
                       't': torch.zeros(2, **f32),                                   # schedule position, twice
                       'sched0': torch.tensor([self.lr.initial, 1.0 - self.wd.initial], **f32),
                       'rates': torch.tensor([self.lr.rate, -self.wd.rate], **f32),
                       'sched': torch.tensor([self.lr.initial, 1.0 - self.wd.initial], **f32),   # lr_t | 1 - wd_t
                       'msum': torch.zeros(1, dtype=torch.float64, device=dev),      # split mode: global sum(mask)
                       'dev_calls': None, 'dev_step': None}
            if dev.type == 'cuda':
                self.opt = _adam(self.bucket.params, lr=self._g['sched'][0], betas=(0.9, 0.9), capturable=True)
            else:
                self.opt = torch.optim.Adam(self.bucket.params, lr=self._g['sched'][0], betas=(0.9, 0.9), eps=1e-7,
                                            foreach=False)
        else:
            self.opt = _adam(self.bucket.params, lr=ft_lr, betas=(0.9, 0.9))             # beta_2 = 0.9 (train.py:310)

    def _tv(self, q, prior, mask, mask_sum):
        return self.trainer.smoothness_loss(torch.cat([prior, mask], -1), q, mask_sum=mask_sum)

    def _fused_loss(self, q, sigma, data, mask, prior, mask_sum, seed=None):
        # equal-size shards (weak scaling): this rank's first global voxel, so ranks draw disjoint Philox counters
        rank = dist.get_rank() if world_size() > 1 else 0
        return self.trainer.fused_elbo(self.layer, q, sigma, data, mask, prior, kl_samples=self.kl_samples,
                                       kl_weight=self.kl_weight, mask_sum=mask_sum, offset=rank * mask.numel(), seed=seed)

    def step(self, data, mask, prior):
        """data [B,X,Y,Z,n_tau] (pre-masked), mask [B,X,Y,Z,1], prior [B,X,Y,Z,5]: this rank's volumes.

        No host synchronisation: the global mask count stays on the device (its all-reduce is enqueued ahead of the
        encoder), the gradient all-reduce runs on a side stream as soon as backward has produced it, and the returned
        statistics are read lazily (LazyStats)."""
        if self.cuda_graph:
            return self._graph_step(data, mask, prior)
        msum = global_mask_sum_device(mask)
        self.bucket.zero_()
        _, q, sigma = self.encoder(data)
        loss, info = self.loss_fn(q, sigma, data, mask, prior, msum)
        tv = self.tv_fn(q, prior, mask, msum)
        total = loss + self.smoothness_weight * tv
        total.backward()
        self._reduce_gradients()
        lr = self.lr(self.step_no)
        for g in self.opt.param_groups:
            g['lr'] = lr
        if self.decay:                                    # tfa AdamW: decoupled decay var -= wd_t * var
            with torch.no_grad():
                self.bucket_params_mul_(1.0 - self.wd(self.step_no))
        self.opt.step()
        self.step_no += 1
        sc = lambda t: t.detach().double().reshape(())                      # noqa: E731
        zero = sc(total) * 0
        stats = torch.stack([sc(total), sc(info['nll']) if 'nll' in info else zero,
                             sc(info['kl']) if 'kl' in info else zero, sc(tv), sc(msum) / max(world_size(), 1)])
        all_reduce_sum_(stats)
        return LazyStats(['loss', 'nll', 'kl', 'smoothness', 'mask_sum'], stats, lr=lr)

    # ---- captured step -------------------------------------------------------------------------------------------
    def static_inputs(self):
        """The (data, mask, prior) buffers the captured step reads, or None before the first step: a loader that
        writes the next batch straight into them saves the three device copies step() otherwise makes."""
        g = self._g
        return None if not g or g['shapes'] is None else (g['data'], g['mask'], g['prior'])

    def _sync_device_scalars(self):
        """Device copies of the Philox call counter and the schedule position follow the host mirrors (they only
        diverge after load_state_dict or when other code drew from the trainer's stream between two steps)."""
        g, tr = self._g, self.trainer
        if g['dev_calls'] != tr._calls:
            g['seed'].fill_(_as_i64((tr._seed + _GOLDEN * tr._calls) & 0xFFFFFFFFFFFFFFFF))
            g['dev_calls'] = tr._calls
        if g['dev_step'] != self.step_no:
            g['t'].fill_(float(self.step_no))
            g['dev_step'] = self.step_no

    def _graph_body(self):
        """One step on the static buffers; identical whether it runs eagerly (warm-up) or under capture."""
        g = self._g
        data, mask, prior = g['data'], g['mask'], g['prior']
        g['seed'].add_(g['golden'])                                   # the key of call number _calls + 1 (uint64 wrap)
        torch.addcmul(g['sched0'], g['t'], g['rates'], out=g['sched'])  # lr_t, 1 - wd_t (LinearSchedule)
        g['t'].add_(1.0)
        msum = global_mask_sum_device(mask)
        self.bucket.zero_()
        _, q, sigma = self.encoder(data)
        loss, info = self.loss_fn(q, sigma, data, mask, prior, msum, seed=g['seed'])
        tv = self.tv_fn(q, prior, mask, msum)
        total = loss + self.smoothness_weight * tv
        total.backward()
        self._reduce_gradients()
        if self.decay:
            with torch.no_grad():
                torch._foreach_mul_(self.bucket.params, g['sched'][1])
        self.opt.step()
        sc = lambda t: t.detach().double().reshape(())                      # noqa: E731
        stats = torch.stack([sc(total), sc(info['nll']), sc(info['kl']), sc(tv), sc(msum) / max(world_size(), 1)])
        all_reduce_sum_(stats)
        return stats

    def _split_front(self):
        """Graph A of the split mode: everything up to and including backward (no collective inside)."""
        g = self._g
        data, mask, prior, msum = g['data'], g['mask'], g['prior'], g['msum']
        g['seed'].add_(g['golden'])
        torch.addcmul(g['sched0'], g['t'], g['rates'], out=g['sched'])
        g['t'].add_(1.0)
        self.bucket.zero_()
        _, q, sigma = self.encoder(data)
        loss, info = self.loss_fn(q, sigma, data, mask, prior, msum, seed=g['seed'])
        tv = self.tv_fn(q, prior, mask, msum)
        total = loss + self.smoothness_weight * tv
        total.backward()
        sc = lambda t: t.detach().double().reshape(())                      # noqa: E731
        return torch.stack([sc(total), sc(info['nll']), sc(info['kl']), sc(tv), sc(msum) / max(world_size(), 1)])

    def _split_back(self):
        """Graph B of the split mode: decoupled weight decay and Adam on the all-reduced gradient."""
        if self.decay:
            with torch.no_grad():
                torch._foreach_mul_(self.bucket.params, self._g['sched'][1])
        self.opt.step()

    def _split_step(self):
        """One step in split mode on the static buffers (eager during warm-up, two graph launches afterwards)."""
        g = self._g
        g['msum'].copy_(global_mask_sum_device(g['mask']))               # eager collective 1 (an input of graph A)
        if g['graph'] is None and g['calls'] >= g['warmup']:
            torch.cuda.synchronize()
            kw = {}
            if world_size() > 1:
                # Other threads of this process make CUDA calls of their own -- the NCCL watchdog polls the events of the
                # collectives issued so far (cudaEventQuery, not allowed while a global-mode capture is under way).  Give
                # it time to retire them (they are complete: the device is idle), and let the capture only police the
                # capturing thread; the kernels the autograd thread adds to the stream are captured either way.
                time.sleep(0.5)
                kw['capture_error_mode'] = 'thread_local'
            front, back = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(front, **kw):
                g['stats'] = self._split_front()
            with torch.cuda.graph(back, pool=front.pool(), **kw):
                self._split_back()
            g['graph'], g['graph_back'] = front, back
        if g['graph'] is not None:
            g['graph'].replay()
            self._reduce_gradients()                                        # eager collective 2
            g['graph_back'].replay()
            stats = g['stats'].clone()
        else:
            stats = self._split_front()
            self._reduce_gradients()
            self._split_back()
        return all_reduce_sum_(stats)                                       # eager collective 3

    def _graph_step(self, data, mask, prior):
        g = self._g
        shapes = (tuple(data.shape), tuple(mask.shape), tuple(prior.shape))
        if g['shapes'] != shapes:                          # first call, or a new batch shape: new buffers, new capture
            g['shapes'], g['graph'], g['calls'] = shapes, None, 0
            g['data'], g['mask'], g['prior'] = (torch.empty_like(t, dtype=torch.float32).contiguous()
                                                for t in (data, mask, prior))
        for name, src in (('data', data), ('mask', mask), ('prior', prior)):
            if src.data_ptr() != g[name].data_ptr():
                g[name].copy_(src, non_blocking=True)
        self._sync_device_scalars()
        lr = self.lr(self.step_no)
        if self.graph_mode == 'split':
            stats = self._split_step()
        elif g['graph'] is None and g['calls'] >= g['warmup']:
            # capture (nothing executes): cuDNN algorithm search, workspaces, the NCCL communicator and the library's
            # lazy state were all set up by the eager warm-up steps
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                g['stats'] = self._graph_body()
            g['graph'] = graph
        if self.graph_mode == 'split':
            pass
        elif g['graph'] is not None:
            g['graph'].replay()
            stats = g['stats'].clone()
        else:
            stats = self._graph_body()
        g['calls'] += 1
        self.step_no += 1
        self.trainer._calls += 1
        g['dev_step'], g['dev_calls'] = self.step_no, self.trainer._calls
        return LazyStats(['loss', 'nll', 'kl', 'smoothness', 'mask_sum'], stats, lr=lr)

    def _reduce_gradients(self):
        """SUM all-reduce of the flat gradient bucket on a side stream (NCCL over NVLink): it is ordered after the
        backward kernels by an event and the optimiser waits for it by an event, so the host thread keeps enqueuing
        (LR schedule, weight decay) while the 585 KB reduction is in flight."""
        if world_size() <= 1:
            return
        if not self.bucket.flat.is_cuda:
            self.bucket.all_reduce_()
            return
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=self.bucket.flat.device)
        cur = torch.cuda.current_stream(self.bucket.flat.device)
        self._comm_stream.wait_stream(cur)
        with torch.cuda.stream(self._comm_stream):
            self.bucket.all_reduce_()
        cur.wait_stream(self._comm_stream)

    def bucket_params_mul_(self, factor):
        torch._foreach_mul_(self.bucket.params, factor)

    # ---- checkpoint / resume (the reference saves pt_model.h5 / final_model.h5 and skips finished phases,
    # train.py:193-202,260-270; here: one torch file with everything the step depends on)
    def state_dict(self):
        return {'encoder': self.encoder.state_dict(), 'optimizer': self.opt.state_dict(), 'step_no': self.step_no,
                'trainer_calls': self.trainer._calls, 'trainer_seed': self.trainer._seed,
                'layer_calls': getattr(self.layer, '_calls', 0)}

    def load_state_dict(self, state):
        self.encoder.load_state_dict(state['encoder'])
        self.opt.load_state_dict(state['optimizer'])
        self.step_no = int(state['step_no'])                      # position on the LR / weight-decay schedule
        self.trainer._calls = int(state.get('trainer_calls', 0))  # Philox call counter: resumed draws continue the stream
        self.trainer._seed = int(state.get('trainer_seed', self.trainer._seed))
        if hasattr(self.layer, '_calls'):
            self.layer._calls = int(state.get('layer_calls', 0))
        self.bucket.realias()
        if self._g is not None:
            # the optimiser's moments are new tensors and its learning rate a loaded copy: re-alias the device schedule
            # and capture again (the device counters follow the host mirrors at the next step)
            for group in self.opt.param_groups:
                group['lr'] = self._g['sched'][0]
            self._g.update(graph=None, calls=0, dev_calls=None, dev_step=None)

    def save(self, path):
        torch.save(self.state_dict(), path)

    def load(self, path, map_location=None):
        self.load_state_dict(torch.load(path, map_location=map_location, weights_only=False))


class StreamingPretrainer:
    """Pre-training on synthetic data (create_and_train_on_synthetic_data, train.py:379-427) without materialising
    the S x S-row dataset (6.25 M rows for the reference's config): every step generates the next slice of the
    shuffled OEF x DBV meshgrid on the device (qbold_generate with the keyed Feistel shuffle, then the noise model),
    feeds it to stream 1 of the encoder in the reference's [-1,10,10,5,n_tau] blocks (train.py:88) and minimises
    synthetic_data_loss.  Ranks walk disjoint row ranges; gradients go through the same flat-bucket all-reduce."""

    def __init__(self, encoder, trainer, params, full_model=True, use_blood=True, uniform_prop=0.1, lr=2e-3,
                 weight_decay=2e-4, batch_blocks=512, seed=1, device=None):
        import ctypes as C
        from . import signals
        self.encoder, self.trainer = encoder, trainer
        self.device = torch.device(device) if device is not None else next(encoder.parameters()).device
        self.layer = signals.SignalGenerationLayer(params, full_model, use_blood, seed=seed)
        gen = torch.Generator(device=self.device).manual_seed(seed)
        S = int(params['sample_size'])
        n_u, n_n = round(S * uniform_prop), round(S * (1.0 - uniform_prop))
        o0, o1 = float(params['oef_start']), float(params['oef_end'])
        d0, d1 = float(params['dbv_start']), float(params['dbv_end'])
        oefs_n = torch.randn(n_n, device=self.device, generator=gen) * float(params['oef_std']) + float(params['oef_mean'])
        self.oefs = torch.cat([torch.rand(n_u, device=self.device, generator=gen) * (o1 - o0) + o0,
                               oefs_n.clamp(o0, o1)]).contiguous()
        self.dbvs = torch.cat([torch.rand(n_u, device=self.device, generator=gen) * (d1 - d0) + d0,
                               signals._truncated_normal(n_n, float(params['dbv_mean']), float(params['dbv_std']), d0, d1,
                                                         gen, self.device)]).contiguous()
        self.total = self.oefs.numel() * self.dbvs.numel()
        self.batch = batch_blocks * 500                                   # 512 blocks of 10x10x5 voxels (train.py:88,103)
        self.seed = seed
        rank = dist.get_rank() if world_size() > 1 else 0
        self.cursor = rank * self.batch
        self.stride = world_size() * self.batch
        self.bucket = FlatGradBucket(encoder.parameters())
        self.opt = _adam(self.bucket.params, lr=lr)
        self.weight_decay = weight_decay
        self._C = C

    def state_dict(self):
        return {'encoder': self.encoder.state_dict(), 'optimizer': self.opt.state_dict(), 'cursor': self.cursor,
                'trainer_calls': self.trainer._calls}

    def load_state_dict(self, state):
        self.encoder.load_state_dict(state['encoder'])
        self.opt.load_state_dict(state['optimizer'])
        self.cursor = int(state['cursor'])                        # position in the shuffled OEF x DBV grid
        self.trainer._calls = int(state.get('trainer_calls', 0))
        self.bucket.realias()

    def next_batch(self):
        """(x [batch, n_tau] noisy signals, y [batch, 3] labels) for this rank's next slice of the shuffled grid."""
        from ._lib import check, dptr, lib, stream_ptr
        C = self._C
        nt = self.layer.n_tau
        first = self.cursor % max(self.total - self.batch, 1)
        self.cursor += self.stride
        x = torch.empty((self.batch, nt), dtype=torch.float32, device=self.device)
        y = torch.empty((self.batch, 3), dtype=torch.float32, device=self.device)
        with _on_device(self.device):
            check(lib().qbold_generate(C.byref(self.layer.params), dptr(self.oefs), self.oefs.numel(), dptr(self.dbvs),
                                       self.dbvs.numel(), None, self.seed, first, self.batch, dptr(x), dptr(y),
                                       stream_ptr(self.device)))
        if self.layer._simulate_noise:
            self.layer.add_noise(x, seed=self.seed ^ 0x5DEECE66D, offset=first, inplace=True)
        return x, y

    def step(self, with_metrics=True):
        x, y = self.next_batch()
        nt = self.layer.n_tau
        self.bucket.zero_()
        blocks = x.reshape(-1, 10, 10, 5, nt)
        fwd = getattr(self.encoder, 'forward_voxelwise', None)           # stream 1 is all the loss sees
        out = fwd(blocks) if fwd is not None else self.encoder(blocks)[0]
        loss = self.trainer.synthetic_data_loss(y, out) / world_size()
        loss.backward()
        self.bucket.all_reduce_()
        if self.weight_decay > 0.0:
            with torch.no_grad():
                torch._foreach_mul_(self.bucket.params, 1.0 - self.weight_decay)
        self.opt.step()
        stat = loss.detach().double().reshape(1)
        all_reduce_sum_(stat)
        if not with_metrics:
            return {'loss': float(stat)}
        with torch.no_grad():                                             # oef / dbv / r2p MSE metrics (model.py:345-374)
            means = self.trainer.calculate_means(out.detach(), None, include_r2p=True, no_samples=20,
                                                 signal_layer=self.layer).reshape(-1, 3)
            mse = ((means - y) ** 2).mean(0)
            vals = torch.cat([stat, mse.double()]).tolist()               # one device -> host read per step
        return {'loss': vals[0], 'oef_mse': vals[1], 'dbv_mse': vals[2], 'r2p_mse': vals[3]}
