#!/usr/bin/env python3
"""BASELINE configs 3-5 on N GPUs (torchrun): amortized-VI training step and whole-volume inference on synthetic
64^3 volumes, voxel/volume shards, NCCL all-reduce of the encoder gradient.  One JSON line per config from rank 0.

    [torchrun --nproc-per-node N] python tools/train_bench.py [--volumes-per-gpu 2] [--steps 10]
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qbold_vi_b200 as qb
from qbold_vi_b200 import distributed as D
from qbold_vi_b200.encoder import create_encoder_from_args


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--volumes-per-gpu', type=int, default=2)
    ap.add_argument('--size', type=int, default=64)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--precision', default='tf32', choices=['tf32', 'fp32', 'bf16'],
                    help='encoder arithmetic: TF32 tensor cores (default), strict fp32, or bf16 autocast; the qBOLD '
                         'kernels are always fp32')
    a = ap.parse_args()
    rank, world, dev = D.init_distributed()
    a.amp = a.precision == 'bf16'
    torch.backends.cuda.matmul.allow_tf32 = a.precision != 'fp32'
    torch.backends.cudnn.allow_tf32 = a.precision != 'fp32'
    torch.backends.cudnn.benchmark = bool(int(os.environ.get("QBOLD_CUDNN_BENCHMARK", "1")))
    args = qb.optimal_arguments()
    cfg = qb.load_system_parameters(qb.config.DEFAULT_CONFIG_PATH)
    cfg['simulate_noise'] = 'False'
    layer = qb.SignalGenerationLayer(cfg, args.full_model, args.use_blood)
    tr = qb.EncoderTrainer(cfg, no_units=args.no_units, no_intermediate_layers=args.no_intermediate_layers,
                           student_t_df=args.student_t_df, initial_im_sigma=args.im_loss_sigma,
                           multi_image_normalisation=args.multi_image_normalisation,
                           channelwise_gating=args.channelwise_gating, use_mvg=args.use_mvg,
                           use_population_prior=args.use_population_prior, predict_log_data=args.predict_log_data,
                           seed=1)
    torch.manual_seed(1)
    enc = create_encoder_from_args(args).to(dev)
    B, S = a.volumes_per_gpu, a.size
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    truth = torch.stack([torch.rand((B, S, S, S), device=dev, generator=g) * 0.5 + 0.15,
                         torch.rand((B, S, S, S), device=dev, generator=g) * 0.1 + 0.01], -1)
    ax = torch.arange(S, device=dev, dtype=torch.float32) - (S - 1) / 2
    r2 = ax[:, None, None] ** 2 + ax[None, :, None] ** 2 + ax[None, None, :] ** 2
    mask = (r2 <= (28.0 * S / 64) ** 2).float()[None, ..., None].expand(B, S, S, S, 1).contiguous()   # centred sphere r=28
    noisy = qb.SignalGenerationLayer(dict(cfg, simulate_noise='True'), True, True, seed=7 + rank)
    data = (noisy(truth.reshape(-1, 2)).reshape(B, S, S, S, 11) * 100.0 * mask).contiguous()
    with torch.no_grad():
        prior = enc(data)[0].clone()
    dp = D.DataParallelTrainer(enc, tr, layer, ft_lr=args.ft_lr, adamw_decay=args.adamw_decay,
                               smoothness_weight=args.smoothness_weight)
    voxels = B * S ** 3

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        for _ in range(3):
            fn()
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        sync()
        t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    stats = {}

    def train_step():
        with torch.autocast('cuda', dtype=torch.bfloat16, enabled=a.amp):
            stats.update(dp.step(data, mask, prior))

    ms_train = timed(train_step, a.steps)
    # where the step time goes on one rank
    def enc_only():
        dp.bucket.zero_()
        with torch.autocast('cuda', dtype=torch.bfloat16, enabled=a.amp):
            _, q, s = enc(data)
        (q.float().sum() + s.float().sum()).backward()
    ms_enc = timed(enc_only, a.steps)
    with torch.no_grad():
        _, q, sigma = enc(data)
    q, sigma = q.contiguous(), sigma.contiguous()
    ms_fused = timed(lambda: tr.fused_elbo(layer, q, sigma, data, mask, prior, kl_samples=70, mask_sum=float(mask.sum())),
                     a.steps)
    ms_inf = timed(lambda: tr.posterior_inference(layer, q, sigma, data, mask, prior, no_samples=64), max(2, a.steps // 3))
    if rank == 0:
        common = {'n_gpus': world, 'volumes_per_gpu': B, 'volume': '%d^3' % S, 'voxels_per_gpu': voxels,
                  'masked_fraction': float(mask.mean()), 'encoder_params': sum(p.numel() for p in enc.parameters()),
                  'encoder_precision': a.precision}
        print(json.dumps(dict(common, config='3/5: VI training step (encoder fwd+bwd, fused ELBO 70-sample KL, TV, '
                              'grad all-reduce, AdamW)', ms_per_step=ms_train,
                              voxel_signals_per_s=world * voxels * 11 / ms_train * 1e3,
                              ms_encoder_fwd_bwd=ms_enc, ms_fused_elbo_kernel=ms_fused, loss=stats.get('loss'))))
        print(json.dumps(dict(common, config='4: posterior inference, 64 samples/voxel (means/variances, likelihood '
                              'map, KL map)', ms_per_volume_batch=ms_inf,
                              voxels_per_s=world * voxels / ms_inf * 1e3)))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
