import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from motion_planning_baselines_b200 import configs
from motion_planning_baselines_b200.fields import CollisionField
from motion_planning_baselines_b200.planners import StochGPMP
from motion_planning_baselines_b200.robots import Robot
dev = dict(device=torch.device('cuda', 0), dtype=torch.float32)
cfg = configs.config('C4')
P, S, H, D = 512, 64, 64, 14
robot = Robot(cfg['robot'], dt=cfg['dt'], tensor_args=dev)
field = CollisionField(cfg['obstacles'], tensor_args=dev)
planner = StochGPMP(robot=robot, n_dof=7, n_support_points=H, num_particles_per_goal=P, opt_iters=1, dt=cfg['dt'],
                    start_state=torch.tensor(cfg['start']).to(**dev), multi_goal_states=torch.tensor(cfg['goal']).to(**dev).unsqueeze(0),
                    collision_fields=[field], tensor_args=dev, num_samples=S, **cfg['params'])
means0 = planner._particle_means.clone()
h_means = means0.cpu().pin_memory()
h_traj = torch.empty(P, H, D).pin_memory()
def run(h2d, d2h, n=20):
    for _ in range(3): planner.optimize(opt_iters=1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(n):
        if h2d: planner._particle_means.copy_(h_means, non_blocking=True)
        traj = planner.optimize(opt_iters=1)
        if d2h: h_traj.copy_(traj, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print(f'h2d={h2d} d2h={d2h}: {e0.elapsed_time(e1)/n:.3f} ms/step (wall {1e3*(time.perf_counter()-t0)/n:.3f})')
run(False, False); run(False, True); run(True, False); run(True, True); run(False, True)
