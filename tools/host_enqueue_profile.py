import os, sys, time, cProfile, pstats, io, torch
sys.path.insert(0, '/root/repo')
exec(open('/root/repo/tools/profile_train_step.py').read().split("from torch.profiler import profile")[0])
torch.cuda.synchronize()
def run(k):
    t0 = time.perf_counter()
    for _ in range(k): dp.step(data, mask, prior)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    return (t1 - t0) / k * 1e3, (t2 - t0) / k * 1e3
print('host enqueue ms/step (3 steps into an empty queue), total ms/step:', run(3))
print('20 steps:', run(20))
pr = cProfile.Profile(); pr.enable()
for _ in range(20): dp.step(data, mask, prior)
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(28); print(s.getvalue()[:6000])
