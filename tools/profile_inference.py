"""Kernel-level breakdown of EncoderTrainer.posterior_inference (BASELINE config 4) on 2 x 64^3 masked volumes."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qbold_vi_b200 as qb
from qbold_vi_b200.encoder import create_encoder_from_args
dev = torch.device('cuda', 0)
cfg = qb.load_system_parameters(qb.config.DEFAULT_CONFIG_PATH); cfg['simulate_noise'] = 'False'
layer = qb.SignalGenerationLayer(cfg, True, True)
tr = qb.EncoderTrainer(cfg, no_units=60, no_intermediate_layers=2, student_t_df=200, multi_image_normalisation=False,
                       channelwise_gating=True, use_mvg=True, use_population_prior=False, predict_log_data=False, seed=1)
torch.manual_seed(1)
enc = create_encoder_from_args(qb.optimal_arguments()).to(dev)
B, S = 2, 64
g = torch.Generator(device=dev).manual_seed(100)
truth = torch.stack([torch.rand((B, S, S, S), device=dev, generator=g) * 0.5 + 0.15,
                     torch.rand((B, S, S, S), device=dev, generator=g) * 0.1 + 0.01], -1)
ax = torch.arange(S, device=dev, dtype=torch.float32) - (S - 1) / 2
r2 = ax[:, None, None] ** 2 + ax[None, :, None] ** 2 + ax[None, None, :] ** 2
mask = (r2 < 28.0 ** 2).float()[None, ..., None].expand(B, S, S, S, 1).contiguous()
data = (layer(truth.reshape(-1, 2)).reshape(B, S, S, S, 11) * 100.0 * mask).contiguous()
with torch.no_grad():
    prior, q, sigma = enc(data)
q, sigma, prior = q.contiguous(), sigma.contiguous(), prior.contiguous()
for _ in range(3):
    tr.posterior_inference(layer, q, sigma, data, mask, prior, no_samples=64)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as p:
    for _ in range(3):
        tr.posterior_inference(layer, q, sigma, data, mask, prior, no_samples=64)
    torch.cuda.synchronize()
print(p.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=70))
