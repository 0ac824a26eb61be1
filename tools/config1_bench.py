#!/usr/bin/env python3
"""BASELINE config 1: `create_synthetic_dataset(params, True, True, 0.0, uniform_prop=0.0)` with sample_size = 1000
(1 M voxels, noise on) through the public API, device-resident result, next to the restated CPU path.

    python tools/config1_bench.py [--reps 5] [--cpu-voxels 100000]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qbold_vi_b200 as qb


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--cpu-voxels', type=int, default=100000)
    a = ap.parse_args()
    cfg = qb.load_system_parameters(qb.config.DEFAULT_CONFIG_PATH)
    cfg['sample_size'] = '1000'
    for _ in range(3):
        x, y = qb.create_synthetic_dataset(cfg, True, True, 0.0, uniform_prop=0.0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(a.reps):
        x, y = qb.create_synthetic_dataset(cfg, True, True, 0.0, uniform_prop=0.0)
    torch.cuda.synchronize()
    sec = (time.perf_counter() - t0) / a.reps
    n = x.shape[0]
    out = {'config': '1: create_synthetic_dataset, sample_size=1000, full model + blood, noise on, 10 noise chunks',
           'voxels': n, 'ms': sec * 1e3, 'voxels_per_s': n / sec, 'voxel_signals_per_s': n * 11 / sec,
           'x_shape': list(x.shape), 'y_shape': list(y.shape), 'finite': bool(torch.isfinite(x).all())}
    # restated CPU path of the same recipe (oracle/torch_port.py forward, all host threads) on a bounded sample
    from oracle import torch_port as tp
    from oracle import qbold_oracle as o
    ph = o.parse_params(o.default_config())
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    rng = np.random.default_rng(0)
    xs = np.stack([rng.uniform(0.05, 0.8, a.cpu_voxels), rng.uniform(0.003, 0.195, a.cpu_voxels)], -1).astype(np.float32)
    with torch.no_grad():
        tp.forward(ph, torch.as_tensor(xs[:8192]))
        t0 = time.perf_counter()
        for i in range(0, a.cpu_voxels, 10000):                   # chunks, as signals.py:282-285 chunks the forward pass
            tp.forward(ph, torch.as_tensor(xs[i:i + 10000]))
    cpu = time.perf_counter() - t0
    out['cpu_port'] = {'voxels': a.cpu_voxels, 'sec': cpu, 'voxels_per_s': a.cpu_voxels / cpu, 'cores': threads}
    out['speedup_vs_cpu_port'] = out['voxels_per_s'] / out['cpu_port']['voxels_per_s']
    print(json.dumps(out), flush=True)


if __name__ == '__main__':
    main()
