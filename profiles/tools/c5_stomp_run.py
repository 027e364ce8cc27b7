"""C5 STOMP at S samples (default 100000), a few iterations -- for ncu launch lists.  usage: c5_stomp_run.py [S]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench_configs
from motion_planning_baselines_b200 import configs
from motion_planning_baselines_b200.costs import CostCollision, CostComposite
from motion_planning_baselines_b200.fields import CollisionField
from motion_planning_baselines_b200.planners import STOMP
from motion_planning_baselines_b200.robots import Robot
dev = dict(device=torch.device('cuda:0'), dtype=torch.float32)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
cfg = configs.config('C5'); prm, Tn, d = cfg['params'], 64, 7
robot = Robot(cfg['robot'], dt=cfg['dt'], tensor_args=dev)
field = CollisionField(cfg['obstacles'], tensor_args=dev)
cost = CostComposite(robot, Tn, [CostCollision(robot, Tn, field=field, sigma_coll=prm['sigma_coll'], tensor_args=dev)], tensor_args=dev)
c1 = configs.config('C1')['params']
planner = STOMP(n_dof=d, n_support_points=Tn, num_particles_per_goal=1, num_samples=S, opt_iters=1, dt=cfg['dt'],
                start_state=torch.tensor(cfg['start']).to(**dev), cost=cost, multi_goal_states=torch.tensor(cfg['goal']).to(**dev).unsqueeze(0),
                temperature=c1['temperature'], step_size=c1['step_size'], sigma_spectral=c1['sigma_spectral'],
                initial_particle_means=bench_configs._straight(cfg, 1, Tn, d, dev), pos_only=False, tensor_args=dev)
for _ in range(3): planner.optimize(opt_iters=1)
torch.cuda.synchronize()
print(bench_configs._time(lambda: planner.optimize(opt_iters=1), 10))
