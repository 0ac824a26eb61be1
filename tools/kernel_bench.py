#!/usr/bin/env python3
"""Per-kernel timing table (CUDA events, 3 warm-ups, inputs larger than L2): one JSON line per kernel.

    python tools/kernel_bench.py [--voxels 4194304] [--reps 5]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qbold_vi_b200 as qb


def timeit(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--voxels', type=int, default=1 << 22)
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--only', default='')
    a = ap.parse_args()
    n, dev = a.voxels, torch.device('cuda', 0)
    cfg = qb.load_system_parameters(qb.config.DEFAULT_CONFIG_PATH)
    cfg['simulate_noise'] = 'False'
    layer = qb.SignalGenerationLayer(cfg, True, True)
    loglin = qb.SignalGenerationLayer(cfg, False, True)
    tr = qb.EncoderTrainer(cfg, student_t_df=200, multi_image_normalisation=False, use_mvg=True,
                           use_population_prior=False, predict_log_data=False, seed=1)
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.rand((n, 2), device=dev, generator=g)
    x[:, 0] = x[:, 0] * 0.8 + 0.04
    x[:, 1] = x[:, 1] * 0.2 + 0.001
    gs = torch.randn((n, 11), device=dev, generator=g)
    q = torch.stack([torch.randn(n, device=dev, generator=g) * 0.7 - 0.3, torch.randn(n, device=dev, generator=g) * 0.6,
                     torch.randn(n, device=dev, generator=g) * 0.7 - 1.2, torch.randn(n, device=dev, generator=g) * 0.6,
                     torch.randn(n, device=dev, generator=g) * 0.8], -1).contiguous()
    prior = (q + 0.3 * torch.randn((n, 5), device=dev, generator=g)).contiguous()
    sigma = torch.exp(torch.randn((n, 11), device=dev, generator=g) * 0.2 - 3.0)
    data = layer(x) * 100.0
    mask = torch.ones(n, device=dev)
    cases = {
        'forward (K1)': lambda: layer(x),
        'forward+VJP (K1b)': lambda: layer.forward_backward(x, gs),
        'forward log-linear': lambda: loglin(x),
        'fused ELBO, 70-sample MC KL (K2)': lambda: tr.fused_elbo(layer, q, sigma, data, mask, prior, kl_samples=70, mask_sum=float(n)),
        'fused ELBO, closed-form KL': lambda: tr.fused_elbo(layer, q, sigma, data, mask, prior, kl_samples=0, mask_sum=float(n)),
        'fused ELBO, no prior': lambda: tr.fused_elbo(layer, q, sigma, data, mask, None, mask_sum=float(n)),
        'kl_loss alone, 70 samples': lambda: tr.kl_loss(torch.cat([prior, mask[:, None]], -1), q, return_mean=False),
        'posterior stats, 64 samples (K4)': lambda: tr.calculate_means(q, None, include_r2p=True, return_stds=True, no_samples=64),
        'reparam sample': lambda: qb.ReparamTrickLayer(tr)((q, None)),
    }
    pred = layer(x)
    y_true = torch.cat([data, mask[:, None]], -1)
    cases['fine_tune_loss_fn alone (k_nll, value + partials)'] = lambda: tr.fine_tune_loss_fn(y_true, torch.cat([pred, sigma], -1))
    side = int((n // 2) ** (1.0 / 3.0) + 1e-6)               # largest cube that fits: 2 x side^3 <= n
    qv = q[:2 * side ** 3].reshape(2, side, side, side, 5).contiguous()
    tv_true = torch.cat([qv, torch.ones_like(qv[..., :1])], -1)
    cases['smoothness TV, 2 x %d^3 (k_smoothness, value + gradient)' % side] = lambda: tr.smoothness_loss(tv_true, qv)
    labels = torch.cat([x, (x[:, :1] * x[:, 1:2]) * 301.74], -1).contiguous()
    cases['pre-training NLL (k_synth_nll, value + gradient)'] = lambda: tr.synthetic_data_loss(labels, q)
    trd = qb.EncoderTrainer(cfg, student_t_df=200, multi_image_normalisation=False, use_mvg=False,
                            use_population_prior=False, predict_log_data=False, seed=1)
    q4, true4 = q[:, :4].contiguous(), torch.cat([prior[:, :4], mask[:, None]], -1).contiguous()
    cases['diagonal KL (k_diag_kl, value + gradient)'] = lambda: trd.kl_loss(true4, q4, return_mean=False)
    from qbold_vi_b200.encoder import create_encoder_from_args
    enc = create_encoder_from_args(qb.optimal_arguments()).to(dev)
    cases['voxel-wise encoder MLP 11-60-60-60-5 (k_encoder_mlp, tcgen05 TF32)'] = lambda: enc.voxelwise_fused(data)
    nl = max(n // 16, 1)
    cases['likelihood map, 64 forward passes per voxel, %d voxels (k_nll_map)' % nl] = lambda: tr.likelihood_map(
        layer, q[:nl], sigma[:nl], data[:nl], None, no_samples=64)
    oefs = torch.rand(2048, device=dev, generator=g) * 0.75 + 0.05
    dbvs = torch.rand(n // 2048, device=dev, generator=g) * 0.19 + 0.003
    cfgn = dict(cfg)
    cfgn['simulate_noise'] = 'True'
    noisy = qb.SignalGenerationLayer(cfgn, True, True, seed=3)
    cases['generate clean (K3, Feistel shuffle)'] = lambda: qb.generate_from_marginals(layer, oefs, dbvs, None, n_chunks=10)
    cases['generate + noise (K3, 10 chunks)'] = lambda: qb.generate_from_marginals(noisy, oefs, dbvs, None, n_chunks=10)
    sig = layer(x)
    cases['noise pass alone (column mean + Philox noise)'] = lambda: noisy.add_noise(sig, inplace=True)
    for name, fn in cases.items():
        if a.only and a.only not in name:
            continue
        ms = timeit(fn, a.reps)
        print(json.dumps({'kernel': name, 'voxels': n, 'ms': round(ms, 4), 'voxels_per_s': n / ms * 1e3,
                          'voxel_signals_per_s': n * 11 / ms * 1e3}), flush=True)


if __name__ == '__main__':
    main()
