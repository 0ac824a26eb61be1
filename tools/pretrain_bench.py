#!/usr/bin/env python3
"""Pre-training step (train.py:379-427 recipe, streaming): generate 256 000 noisy voxels on the device, stream 1 of the
encoder, logit-normal NLL kernel, backward, AdamW -- ms per step on one GPU.

    python tools/pretrain_bench.py [--steps 30]
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qbold_vi_b200 as qb
from qbold_vi_b200 import distributed as D
from qbold_vi_b200.encoder import create_encoder_from_args


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=30)
    a = ap.parse_args()
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = True
    dev = torch.device('cuda', 0)
    args = qb.optimal_arguments()
    cfg = qb.load_system_parameters(qb.config.DEFAULT_CONFIG_PATH)
    torch.manual_seed(1)
    enc = create_encoder_from_args(args).to(dev)
    tr = qb.EncoderTrainer(cfg, no_units=60, no_intermediate_layers=2, student_t_df=200, multi_image_normalisation=False,
                           channelwise_gating=True, use_mvg=True, use_population_prior=False, predict_log_data=False, seed=1)
    pt = D.StreamingPretrainer(enc, tr, cfg, True, True, uniform_prop=0.0, lr=2e-3, device=dev)
    for _ in range(3):
        stats = pt.step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        stats = pt.step()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / a.steps * 1e3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        pt.next_batch()
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({'config': 'pre-training step: streaming generation + noise, stream-1 encoder, NLL kernel, AdamW',
                      'voxels_per_step': pt.batch, 'ms_per_step': ms, 'voxels_per_s': pt.batch / ms * 1e3,
                      'ms_generate_batch': e0.elapsed_time(e1) / 10, 'loss': stats['loss']}))


if __name__ == '__main__':
    main()
