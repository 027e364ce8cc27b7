import os, sys, torch
sys.path.insert(0, '.')
from motion_planning_baselines_b200 import _lib
P, S, M = 512, 64, 896
dev = dict(device='cuda', dtype=torch.float32)
L = torch.tril(torch.randn(M, M, **dev)) / 30
split = torch.empty(2, M, M, **dev)
lib = _lib.lib()
_lib.check(lib.mpb_split_tf32(_lib.ptr(L), _lib.ptr(split[0]), _lib.ptr(split[1]), M * M, _lib.stream_ptr()))
mu = torch.randn(P, M, **dev); eps = [torch.randn(S, P, M, **dev) for _ in range(4)]; x = torch.empty(P, S, M, **dev)
def run(n):
    for i in range(n):
        _lib.check(lib.mpb_sample_gp_tc(_lib.ptr(split[0]), _lib.ptr(split[1]), _lib.ptr(mu), _lib.ptr(eps[i % 4]), _lib.ptr(x), P, S, M, _lib.stream_ptr()))
run(5); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(50); e1.record(); torch.cuda.synchronize()
print('MPB_TC_DEBUG', os.environ.get('MPB_TC_DEBUG', '0'), 'ms', e0.elapsed_time(e1) / 50)
