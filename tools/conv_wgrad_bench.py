#!/usr/bin/env python3
"""qbold_conv_wgrad (tcgen05) against cuDNN's weight-gradient kernel on the training shape (2 x 64^3 voxels, 60 channels)."""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qbold_vi_b200 as qb
from qbold_vi_b200._lib import check, dptr, lib, stream_ptr
torch.backends.cudnn.allow_tf32 = True; torch.backends.cudnn.benchmark = True
dev = torch.device('cuda', 0)
bz, nx, ny, c = 128, 64, 64, 60
x = torch.randn(bz, nx, ny, c, device=dev); g = torch.randn(bz, nx, ny, c, device=dev)
w = torch.randn(c, c, 3, 3, device=dev).contiguous(memory_format=torch.channels_last)
dw = torch.empty(c, c, 3, 3, device=dev); ws = torch.empty(int(lib().qbold_conv_wgrad_workspace_floats()), device=dev)
status = torch.zeros(1, dtype=torch.int32, device=dev)
def ours():
    check(lib().qbold_conv_wgrad(dptr(g.reshape(-1, c)), c, dptr(x.reshape(-1, c)), c, bz, nx, ny, dptr(dw), 0, dptr(ws), dptr(status, torch.int32), stream_ptr(dev)))
xi, gi = x.permute(0, 3, 1, 2), g.permute(0, 3, 1, 2)
def cudnn():
    return torch.ops.aten.convolution_backward(gi, xi, w, None, (1, 1), (1, 1), (1, 1), False, (0, 0), 1, (False, True, False))
def timed(fn, k=20):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k * 1e3
t_o, t_c = timed(ours), timed(cudnn)
ref = cudnn()[1]
err = float((dw - ref).abs().max() / ref.abs().max())
print(json.dumps({'shape': [bz, nx, ny, c], 'us_tcgen05': t_o, 'us_cudnn': t_c, 'rel_diff_vs_cudnn': err, 'status': int(status.item()),
                  'tflops_tcgen05': 2 * bz * nx * ny * c * c * 9 / t_o / 1e6}))
