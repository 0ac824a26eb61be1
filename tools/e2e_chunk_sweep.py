import os, sys, time, ctypes as C, json, torch
sys.path.insert(0, '/root/repo')
import qbold_vi_b200 as qb
from qbold_vi_b200 import _lib
dev = torch.device('cuda', 0)
cfg = qb.load_system_parameters(qb.config.DEFAULT_CONFIG_PATH); cfg['simulate_noise'] = 'False'
layer = qb.SignalGenerationLayer(cfg, True, True)
n = 1 << 24
hx = (torch.rand(n, 2) * torch.tensor([0.8, 0.2]) + torch.tensor([0.04, 0.001])).pin_memory()
hg = torch.randn(n, 11).pin_memory()
hs = torch.empty(n, 11).pin_memory(); hgr = torch.empty(n, 2).pin_memory()
lib = _lib.lib()
def run():
    _lib.check(lib.qbold_forward_backward_host(C.byref(layer.params), hx.data_ptr(), hg.data_ptr(), n, hs.data_ptr(), hgr.data_ptr()))
run(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3): run()
t = (time.perf_counter() - t0) / 3
gb = (C.c_double * 2)()
_lib.check(lib.qbold_host_copy_ceiling(hg.data_ptr(), hs.data_ptr(), hg.numel() * 4, 3, gb))
print(json.dumps({'chunk_log2': os.environ.get('QBOLD_HOST_CHUNK_LOG2', '17'), 'ms': t * 1e3, 'gbs_per_dir': n * 52 / t / 1e9, 'ceiling': gb[0], 'frac': n * 52 / t / 1e9 / gb[0]}))
