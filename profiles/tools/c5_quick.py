"""C5 MPPI / STOMP at 1e5 (and MPPI 1e6) samples: ms per iteration.  usage: c5_quick.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench_configs
dev = dict(device=torch.device('cuda:0'), dtype=torch.float32)
r = bench_configs.bench_c5(dev, Ns=(100000, 1000000), stomp_Ns=(100000,), iters=10)
print([(x['config'][-5:], x['shape'][:8], round(x['ms_per_iter'], 4)) for x in r])
