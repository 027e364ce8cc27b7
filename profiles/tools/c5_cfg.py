"""C5 (Panda, table + shelf boxes) MPPI / STOMP at 1e5 samples under the packed-K2 launch shapes (MPB_K2_CFG).  usage: c5_cfg.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench_configs
dev = dict(device=torch.device('cuda:0'), dtype=torch.float32)
for cfg in ('8x2', '10x2', '6x3'):
    os.environ['MPB_K2_CFG'] = cfg
    r = bench_configs.bench_c5(dev, Ns=(100000,), stomp_Ns=(100000,), iters=10)
    print(cfg, [(x['config'][-5:], round(x['ms_per_iter'], 4)) for x in r])
