"""Kernel-level breakdown (torch profiler) of the fine-tuning step on 2 x 64^3 unmasked volumes."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qbold_vi_b200 as qb
from qbold_vi_b200 import distributed as D
from qbold_vi_b200.encoder import create_encoder_from_args
torch.backends.cuda.matmul.allow_tf32 = True; torch.backends.cudnn.allow_tf32 = True; torch.backends.cudnn.benchmark = True
dev = torch.device('cuda', 0)
args = qb.optimal_arguments()
cfg = qb.load_system_parameters(qb.config.DEFAULT_CONFIG_PATH); cfg['simulate_noise'] = 'False'
layer = qb.SignalGenerationLayer(cfg, True, True)
tr = qb.EncoderTrainer(cfg, no_units=60, no_intermediate_layers=2, student_t_df=200, multi_image_normalisation=False, channelwise_gating=True, use_mvg=True, use_population_prior=False, predict_log_data=False, seed=1)
torch.manual_seed(1)
enc = create_encoder_from_args(args).to(dev)
B, S = 2, 64
g = torch.Generator(device=dev).manual_seed(100)
truth = torch.stack([torch.rand((B, S, S, S), device=dev, generator=g) * 0.5 + 0.15, torch.rand((B, S, S, S), device=dev, generator=g) * 0.1 + 0.01], -1)
mask = torch.ones(B, S, S, S, 1, device=dev)
data = (layer(truth.reshape(-1, 2)).reshape(B, S, S, S, 11) * 100.0).contiguous()
with torch.no_grad(): prior = enc(data)[0].clone()
dp = D.DataParallelTrainer(enc, tr, layer)
for _ in range(3): dp.step(data, mask, prior)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as p:
    for _ in range(3): dp.step(data, mask, prior)
    torch.cuda.synchronize()
print(p.key_averages().table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=70))
# GPU-busy time per step against the wall time of a step (launch gaps show up as the difference)
import time
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): dp.step(data, mask, prior)
torch.cuda.synchronize(); wall = (time.perf_counter() - t0) / 10 * 1e3
busy = sum(e.self_device_time_total for e in p.key_averages()) / 3 / 1e3
print('wall ms/step %.3f   GPU-busy ms/step (sum of kernel times) %.3f' % (wall, busy))
