"""K2 time vs batch size at the C4 shape: the intercept is the per-CTA prologue (staging the robot / field tables, cull table,
link reach test) + one trajectory per warp.  usage: k2_prologue.py"""
import os, sys, ctypes as C
import torch
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
from motion_planning_baselines_b200 import _lib
dev = dict(device=torch.device('cuda:0'), dtype=torch.float32)
from test_gpu_bench_shape import build
cfg, sig, pl = build('C4', dev)
pl.optimize(opt_iters=1)
lib, st = _lib.lib(), _lib.stream_ptr()
gp, fields, nf, _ = pl.cost._build()
P, S, H = pl.num_particles, pl.num_samples, pl.n_support_points
x = pl.state_samples.view(P * S, H, -1)
for B in (296, 2960, 5920, 11840, 32768):
    def run():
        _lib.check(lib.mpb_cost_eval(_lib.ptr(x), B, H, C.byref(pl.robot.desc), fields, nf, C.byref(gp), _lib.ptr(pl._is_vec), S,
                                     pl.temperature, _lib.ptr(pl.costs), None, _lib.ptr(pl.free_flags), st))
    for _ in range(3): run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): run()
    e1.record(); torch.cuda.synchronize()
    print(f'B = {B:6d}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per launch')
