#!/usr/bin/env python3
"""qbold_dense_tma against the cuBLAS calls it replaces in the encoder's training step (524 288 x 60 -> 60, TF32)."""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qbold_vi_b200 as qb
from qbold_vi_b200._lib import check, dptr, lib, stream_ptr
torch.backends.cuda.matmul.allow_tf32 = True
dev = torch.device('cuda', 0)
n, c = 2 * 64 ** 3, 60
x = torch.randn(n, c, device=dev); w = torch.randn(c, c, device=dev) * 0.2; b = torch.randn(c, device=dev)
y = torch.empty(n, c, device=dev); acc = torch.randn(n, c, device=dev)
status = torch.zeros(1, dtype=torch.int32, device=dev)
def tma(transpose, relu, add):
    check(lib().qbold_dense_tma(dptr(x), dptr(w), None if transpose else dptr(b), dptr(acc) if add else None, c, c, transpose, relu, n,
                                dptr(acc) if add else dptr(y), dptr(status, torch.int32), stream_ptr(dev)))
def timed(fn, k=20):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k * 1e3
rows = {
    'forward bias+relu': (lambda: tma(0, 1, 0), lambda: torch._addmm_activation(b, x, w.t(), use_gelu=False)),
    'forward bias': (lambda: tma(0, 0, 0), lambda: torch.addmm(b, x, w.t())),
    'input gradient': (lambda: tma(1, 0, 0), lambda: x @ w),
    'input gradient, in-place accumulate': (lambda: tma(1, 0, 1), lambda: acc.addmm_(x, w)),
}
for name, (ours, cublas) in rows.items():
    t_o, t_c = timed(ours), timed(cublas)
    moved = n * c * 4 * (3 if 'accumulate' in name else 2)
    print(json.dumps({'op': name, 'us_tma_tcgen05': round(t_o, 1), 'us_cublas': round(t_c, 1), 'gbs_tma': round(moved / t_o / 1e3, 1),
                      'status': int(status.item())}))
