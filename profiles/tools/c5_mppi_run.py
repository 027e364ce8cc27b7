"""C5 MPPI at N control samples (default 100000), a few iterations -- for ncu launch lists.  usage: c5_mppi_run.py [N]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench_configs
dev = dict(device=torch.device('cuda:0'), dtype=torch.float32)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
print(bench_configs.bench_c5(dev, Ns=(N,), stomp_Ns=(), iters=5))
