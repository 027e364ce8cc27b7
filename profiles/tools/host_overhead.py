"""Host-side (Python + ctypes + launch) time of one planner.optimize(opt_iters=1) call at C4 vs the device time of the step.
usage: host_overhead.py"""
import os, sys, time
import torch
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
dev = dict(device=torch.device('cuda:0'), dtype=torch.float32)
from test_gpu_bench_shape import build
cfg, sig, pl = build('C4', dev)
for _ in range(5):
    pl.optimize(opt_iters=1)
torch.cuda.synchronize()
for n in (10, 20):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        pl.optimize(opt_iters=1)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f'{n} calls: host enqueue {1e6 * (t1 - t0) / n:.1f} us/call, wall incl. drain {1e6 * (t2 - t0) / n:.1f} us/call')
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(20):
    pl.optimize(opt_iters=1)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('cumulative').print_stats(14)
