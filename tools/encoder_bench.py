#!/usr/bin/env python3
"""Voxel-wise encoder branch (stream 1): the tcgen05 kernel vs the same layers through torch (cuBLAS TF32).

    python tools/encoder_bench.py [--voxels 2097152] [--reps 10]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qbold_vi_b200 as qb
from qbold_vi_b200.encoder import create_encoder_from_args


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
    ap.add_argument('--voxels', type=int, default=1 << 21)
    ap.add_argument('--reps', type=int, default=10)
    a = ap.parse_args()
    dev = torch.device('cuda', 0)
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.manual_seed(0)
    enc = create_encoder_from_args(qb.optimal_arguments()).to(dev)
    n = a.voxels
    data = torch.rand((n, 11), device=dev) * 150.0 + 20.0

    def torch_stream1():
        with torch.no_grad():
            h = enc.act(enc.first(enc.normalise_data(data)))
            for b in enc.blocks:
                h = enc.act(b.pointwise(h))
            return enc.final(h)

    ms_t = timeit(torch_stream1, a.reps)
    ms_k = timeit(lambda: enc.voxelwise_fused(data), a.reps)
    err = float((enc.voxelwise_fused(data) - torch_stream1()).abs().max())
    flops = 2.0 * (11 * 60 + 2 * 60 * 60 + 60 * 5)
    for name, ms in (('torch (normalise + 4 cuBLAS TF32 GEMMs + ReLUs)', ms_t), ('k_encoder_mlp (tcgen05, one launch)', ms_k)):
        print(json.dumps({'path': name, 'voxels': n, 'ms': round(ms, 4), 'voxels_per_s': n / ms * 1e3,
                          'hbm_gbs_alg': n * 64 / ms / 1e6, 'useful_tflops': n * flops / ms / 1e9,
                          'max_abs_diff_between_paths': err}), flush=True)


if __name__ == '__main__':
    main()
