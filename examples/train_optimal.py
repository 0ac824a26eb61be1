#!/usr/bin/env python3
"""End-to-end run of the reference's train.py flow on SYNTHETIC volumes with optimal.yaml, on 1..N GPUs:

    pre-train stream 1 on streamed synthetic signals  (create_and_train_on_synthetic_data, train.py:379-427)
 -> priors from the pre-trained model                 (prepare_dataset, train.py:26-31)
 -> fine-tune with the fused ELBO + TV                (train_full_model, train.py:285-376)
 -> whole-volume posterior inference                  (save_predictions, model.py:772-887, without NIfTI I/O)

    [torchrun --nproc-per-node N] python examples/train_optimal.py [--pt-steps 150] [--ft-steps 60] [--size 32]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qbold_vi_b200 as qb
from qbold_vi_b200 import distributed as D
from qbold_vi_b200.encoder import create_encoder_from_args


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--pt-steps', type=int, default=150)
    ap.add_argument('--ft-steps', type=int, default=60)
    ap.add_argument('--size', type=int, default=32)
    ap.add_argument('--volumes-per-gpu', type=int, default=2)
    ap.add_argument('--save-directory', default=None,
                    help='checkpoints: pt_model.pt / final_model.pt; finished phases are skipped (train.py:193-202,260-270)')
    ap.add_argument('--cuda-graph', default='off', choices=['off', 'full', 'split'],
                    help="fine-tuning step replayed from a CUDA graph: 'full' (one graph, measured on one GPU) or 'split' "
                         "(collectives outside the capture, for several GPUs)")
    a = ap.parse_args()
    rank, world, dev = D.init_distributed()
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = True
    torch.backends.cudnn.benchmark = True          # let cuDNN pick the 3x3x1 kernels (fixed shapes per run)
    args = qb.optimal_arguments()                                   # configurations/optimal.yaml over train.py defaults
    params = qb.load_system_parameters(qb.config.DEFAULT_CONFIG_PATH)
    params['sample_size'] = '1000'
    trainer = qb.EncoderTrainer(system_params=params, no_units=args.no_units, use_layer_norm=args.use_layer_norm,
                                dropout_rate=args.dropout_rate, no_intermediate_layers=args.no_intermediate_layers,
                                student_t_df=args.student_t_df, initial_im_sigma=args.im_loss_sigma,
                                activation_type=args.activation, multi_image_normalisation=args.multi_image_normalisation,
                                channelwise_gating=args.channelwise_gating, infer_inv_gamma=False,
                                use_population_prior=args.use_population_prior, use_mvg=args.use_mvg,
                                predict_log_data=args.predict_log_data, seed=1)
    torch.manual_seed(1)
    model = create_encoder_from_args(args).to(dev)

    # ---- pre-training on streamed synthetic data (noise on, as the INI says)
    pt = D.StreamingPretrainer(model, trainer, params, args.full_model, args.use_blood, uniform_prop=args.uniform_prop,
                               lr=args.pt_lr, weight_decay=args.pt_adamw_decay, batch_blocks=128, seed=1 + rank, device=dev)
    log = []
    ckpt = (lambda name: os.path.join(a.save_directory, name)) if a.save_directory else None
    if ckpt:
        os.makedirs(a.save_directory, exist_ok=True)
    if ckpt and os.path.exists(ckpt('pt_model.pt')):            # the reference loads pt_model.h5 instead of pre-training again
        pt.load_state_dict(torch.load(ckpt('pt_model.pt'), map_location=dev, weights_only=False))
        log.append(('pt', 'resumed', {'cursor': pt.cursor}))
    else:
        for i in range(a.pt_steps):
            st = pt.step()
            if i % 25 == 0 or i == a.pt_steps - 1:
                log.append(('pt', i, st))
        if ckpt and rank == 0:
            torch.save(pt.state_dict(), ckpt('pt_model.pt'))

    # ---- synthetic "scans": smooth OEF/DBV maps inside a sphere, forward model + noise
    S, B = a.size, a.volumes_per_gpu
    g = torch.Generator(device=dev).manual_seed(50 + rank)
    ax = torch.linspace(-1, 1, S, device=dev)
    xx, yy, zz = torch.meshgrid(ax, ax, ax, indexing='ij')
    oef = 0.35 + 0.12 * torch.sin(2.5 * xx) * torch.cos(2.0 * yy) + 0.03 * zz
    dbv = 0.03 + 0.015 * torch.cos(3.0 * xx + 1.0) * torch.sin(2.0 * zz)
    truth = torch.stack([oef, dbv], -1).expand(B, S, S, S, 2).contiguous()
    mask = ((xx ** 2 + yy ** 2 + zz ** 2) <= 0.85 ** 2).float()[None, ..., None].expand(B, S, S, S, 1).contiguous()
    noisy = qb.SignalGenerationLayer(params, args.full_model, args.use_blood, seed=9 + rank)     # simulate_noise = True
    data = (noisy(truth.reshape(-1, 2)).reshape(B, S, S, S, -1) * 100.0 * mask).contiguous()

    # ---- fine-tuning: priors from the pre-trained model, noise off in the forward model (train.py:256)
    params['simulate_noise'] = 'False'
    sig_gen_layer = qb.SignalGenerationLayer(params, args.full_model, args.use_blood)
    with torch.no_grad():
        prior = model(data)[0][..., :5].contiguous()
    ft = D.DataParallelTrainer(model, trainer, sig_gen_layer, ft_lr=args.ft_lr, adamw_decay=args.adamw_decay,
                               smoothness_weight=args.smoothness_weight, kl_weight=1.0,
                               cuda_graph=False if a.cuda_graph == 'off' else a.cuda_graph)
    if ckpt and os.path.exists(ckpt('final_model.pt')):         # resume: weights, Adam moments, schedule position, RNG counters
        ft.load(ckpt('final_model.pt'), map_location=dev)
        log.append(('ft', 'resumed', {'step_no': ft.step_no}))
    for i in range(ft.step_no, a.ft_steps):
        st = ft.step(data, mask, prior)
        if i % 10 == 0 or i == a.ft_steps - 1:
            log.append(('ft', i, st.as_dict()))                 # reading the lazy statistics is the only host sync
    if ckpt and rank == 0:
        ft.save(ckpt('final_model.pt'))

    # ---- inference: 64 posterior samples per voxel
    with torch.no_grad():
        _, q, sigma = model(data)
    res = trainer.posterior_inference(sig_gen_layer, q.contiguous(), sigma.contiguous(), data, mask, prior, no_samples=64)
    m = mask[..., 0] > 0
    err = (res['means'][..., :2] - truth).abs()[m].mean(0)
    if rank == 0:
        for phase, i, st in log:
            print(json.dumps({'phase': phase, 'step': i, **{k: (round(v, 5) if isinstance(v, float) else v) for k, v in st.items()}}))
        print(json.dumps({'phase': 'inference', 'mean_abs_err_oef': float(err[0]), 'mean_abs_err_dbv': float(err[1]),
                          'mean_likelihood': float(res['likelihood'][m].mean()), 'mean_kl': float(res['kl'][m].mean()),
                          'n_gpus': world}))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
