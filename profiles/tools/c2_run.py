import sys, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench_configs
dev = dict(device=torch.device('cuda:0'), dtype=torch.float32)
print(bench_configs.bench_c2(dev, iters=20))
