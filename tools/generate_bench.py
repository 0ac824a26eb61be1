#!/usr/bin/env python3
"""BASELINE config 1/5: streaming synthetic generation (create_synthetic_dataset recipe with in-kernel Feistel
shuffle + Philox noise) of V voxels in shards of <= 16 M per launch, per GPU; one JSON line from rank 0.

    [torchrun --nproc-per-node N] python tools/generate_bench.py [--voxels 1073741824]
"""
import argparse
import ctypes as C
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qbold_vi_b200 as qb
from qbold_vi_b200 import distributed as D
from qbold_vi_b200._lib import check, dptr, lib, stream_ptr


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--voxels', type=int, default=1 << 30)
    ap.add_argument('--shard', type=int, default=1 << 24)
    a = ap.parse_args()
    rank, world, dev = D.init_distributed()
    cfg = qb.load_system_parameters(qb.config.DEFAULT_CONFIG_PATH)        # simulate_noise = True
    layer = qb.SignalGenerationLayer(cfg, True, True, seed=1)
    side = int(round(a.voxels ** 0.5))
    total = side * side
    gen = torch.Generator(device=dev).manual_seed(1)
    oefs = (torch.randn(side, device=dev, generator=gen) * 0.2 + 0.4).clamp(0.05, 0.8).contiguous()      # signals.py:258-259
    from qbold_vi_b200.signals import _truncated_normal
    dbvs = _truncated_normal(side, 0.025, 0.02, 0.003, 0.195, gen, dev).contiguous()                       # signals.py:265-267
    lo, hi = D.shard_range(total, rank, world)                           # this rank's rows of the shuffled meshgrid
    x = torch.empty((a.shard, 11), device=dev)
    y = torch.empty((a.shard, 3), device=dev)
    mean = torch.empty(11, device=dev)
    scratch = torch.empty(22, dtype=torch.float64, device=dev)
    P, st, L = C.byref(layer.params), stream_ptr(dev), lib()

    def run():
        for first in range(lo, hi, a.shard):
            m = min(a.shard, hi - first)
            check(L.qbold_generate(P, dptr(oefs), side, dptr(dbvs), side, None, 1, first, m, dptr(x), dptr(y), st))
            check(L.qbold_column_mean(dptr(x), m, 11, dptr(mean), dptr(scratch, torch.float64), st))
            check(L.qbold_add_noise(P, dptr(x), m, dptr(mean), None, None, 2, first, st))

    # warm-up on a small slice, then the full shard
    hi_saved, hi = hi, min(hi, lo + a.shard)
    run()
    hi = hi_saved
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(t)
        print(json.dumps({'config': '1/5: synthetic generation, %d voxels (%d x %d meshgrid, Feistel shuffle, Philox noise '
                                    'per 16M-voxel chunk)' % (total, side, side), 'n_gpus': world, 'ms': ms,
                          'voxels_per_s': total / ms * 1e3, 'voxel_signals_per_s': total * 11 / ms * 1e3,
                          'sample_row': x[0].tolist(), 'finite': bool(torch.isfinite(x).all())}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
