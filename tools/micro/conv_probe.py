"""Probe: cuDNN 3x3x1 conv (channels-last 3d, TF32) at C=60 vs C=64, fused bias+ReLU op, Dense GEMMs at 60 vs 64."""
import torch, time
torch.backends.cuda.matmul.allow_tf32 = True; torch.backends.cudnn.allow_tf32 = True; torch.backends.cudnn.benchmark = True
dev = torch.device('cuda', 0)
def timed(fn, k=20):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k * 1e3
B, S = 2, 64
for C in (60, 64):
    x = torch.randn(B, S, S, S, C, device=dev).permute(0, 4, 1, 2, 3)          # channels_last_3d view
    conv = torch.nn.Conv3d(C, C, (3, 3, 1), padding=(1, 1, 0)).to(dev)
    conv.weight.data = conv.weight.data.contiguous(memory_format=torch.channels_last_3d)
    w, b = conv.weight, conv.bias
    t_f = timed(lambda: torch.nn.functional.conv3d(x, w, None, padding=(1, 1, 0)))
    t_fb = timed(lambda: torch.nn.functional.conv3d(x, w, b, padding=(1, 1, 0)))
    try:
        t_fr = timed(lambda: torch.cudnn_convolution_relu(x, w, b, (1, 1, 1), (1, 1, 0), (1, 1, 1), 1))
        y1 = torch.cudnn_convolution_relu(x, w, b, (1, 1, 1), (1, 1, 0), (1, 1, 1), 1)
        y2 = torch.relu(torch.nn.functional.conv3d(x, w, b, padding=(1, 1, 0)))
        err = float((y1 - y2).abs().max()); cl = y1.is_contiguous(memory_format=torch.channels_last_3d)
    except Exception as e:
        t_fr, err, cl = None, str(e)[:100], None
    g = torch.randn_like(conv(x))
    xr = x.detach().requires_grad_(True)
    def bwd():
        y = torch.nn.functional.conv3d(xr, w, None, padding=(1, 1, 0))
        return torch.autograd.grad(y, [xr, w], g)
    t_fwdbwd = timed(bwd)
    gi = lambda: torch.ops.aten.convolution_backward(g, x, w, None, (1, 1, 1), (1, 1, 0), (1, 1, 1), False, (0, 0, 0), 1, (True, False, False))
    gw = lambda: torch.ops.aten.convolution_backward(g, x, w, None, (1, 1, 1), (1, 1, 0), (1, 1, 1), False, (0, 0, 0), 1, (False, True, False))
    print('C=%d conv fwd %.0f us, fwd+bias %.0f us, fused bias+relu %s us (maxdiff %s, channels_last %s), dgrad %.0f us, wgrad %.0f us'
          % (C, t_f, t_fb, t_fr, err, cl, timed(gi), timed(gw)))
    n = B * S ** 3
    a = torch.randn(n, C, device=dev); W = torch.randn(C, C, device=dev); bb = torch.randn(C, device=dev)
    print('   dense addmm %.0f us, addmm+relu %.0f us, mm %.0f us, relu %.0f us, add_ %.0f us'
          % (timed(lambda: torch.addmm(bb, a, W.t())), timed(lambda: torch._addmm_activation(bb, a, W.t(), use_gelu=False)),
             timed(lambda: a @ W), timed(lambda: torch.relu(a)), timed(lambda: a.add_(a))))
# bf16 conv for reference
x = torch.randn(B, S, S, S, 64, device=dev, dtype=torch.bfloat16).permute(0, 4, 1, 2, 3)
w = torch.randn(64, 64, 3, 3, 1, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
print('bf16 C=64 conv fwd %.0f us' % timed(lambda: torch.nn.functional.conv3d(x, w, None, padding=(1, 1, 0))))
# 2-D formulation: fold z into the batch: [B*Z, C, X, Y] channels_last 2d conv 3x3
for C in (60, 64):
    x2 = torch.randn(B * S, S, S, C, device=dev).permute(0, 3, 1, 2)
    w2 = torch.randn(C, C, 3, 3, device=dev).contiguous(memory_format=torch.channels_last)
    print('2d conv (z folded into batch) C=%d fwd %.0f us' % (C, timed(lambda: torch.nn.functional.conv2d(x2, w2, None, padding=1))))
