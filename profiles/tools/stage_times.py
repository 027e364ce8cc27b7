"""Per-stage CUDA-event times of one Stoch-GPMP iteration (K1 [+ Sigma^-1 mu], mat-vec, K2, K3) for a config.  usage: stage_times.py C3|C4"""
import os, sys
import torch
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
dev = dict(device=torch.device('cuda:0'), dtype=torch.float32)
from test_gpu_bench_shape import build
name = sys.argv[1] if len(sys.argv) > 1 else 'C3'
cfg, sig, pl = build(name, dev)
n = 30
for i in range(3): pl.step_staged(None)
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(n)]
torch.cuda.synchronize()
for i in range(n): pl.step_staged(None, events=ev[i])
torch.cuda.synchronize()
st = [sum(ev[i][j].elapsed_time(ev[i][j + 1]) for i in range(n)) / n for j in range(4)]
print(name, 'K1 / mat-vec / K2 / K3 ms:', [round(x, 4) for x in st], 'sum', round(sum(st), 4))
