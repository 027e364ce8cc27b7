import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from motion_planning_baselines_b200 import configs
from motion_planning_baselines_b200.costs import build_gpmp2_cost_composite
from motion_planning_baselines_b200.fields import CollisionField
from motion_planning_baselines_b200.robots import Robot
from oracle.build import TA, oracle_field, oracle_robot
from oracle.costs import CostSpec
dev = dict(device=torch.device('cuda:0'), dtype=torch.float32)
cfg_name, P, S, H = 'C1', 5, 16, 64
cfg = configs.config(cfg_name)
model, obst = cfg['robot'], cfg['obstacles']
d = model.q_dim
first_cpu = first_gpu = None
for rep in range(int(sys.argv[1])):
    gen = torch.Generator().manual_seed(1234 + H)
    start, goal = torch.tensor(cfg['start']), torch.tensor(cfg['goal'])
    w = torch.linspace(0, 1, H).view(1, H, 1)
    line = start * (1 - w) + goal * w
    x = torch.zeros(P * S, H, 2 * d)
    x[..., :d] = line + 0.15 * torch.randn(P * S, H, d, generator=gen).cumsum(1) / np.sqrt(H) + 0.05 * torch.randn(P * S, 1, d, generator=gen)
    x[..., d:] = 0.5 * torch.randn(P * S, H, d, generator=gen)
    sig = dict(sigma_start=1e-2, sigma_gp=1.0, sigma_goal_prior=1e-2, sigma_coll=1e-1)
    spec = CostSpec(oracle_robot(model, cfg['dt']), H, cfg['dt'], start, goal, [oracle_field(obst, model)], tensor_args=TA, **sig)
    ref_terms = torch.stack(spec.terms(x))
    ref_free = spec.collision_free(x)
    robot = Robot(model, dt=cfg['dt'], tensor_args=dev)
    comp = build_gpmp2_cost_composite(robot=robot, n_support_points=H, dt=cfg['dt'], start_state=start.to(**dev),
                                      multi_goal_states=goal.to(**dev).unsqueeze(0), num_particles_per_goal=P,
                                      collision_fields=[CollisionField(obst, tensor_args=dev)], num_samples=S, tensor_args=dev, **sig)
    xg = x.to(**dev)
    terms, _ = comp.eval(xg, return_invidual_costs_and_weights=True)
    tg = torch.stack(terms).cpu()
    if first_cpu is None:
        first_cpu, first_gpu, first_x = ref_terms.clone(), tg.clone(), x.clone()
    else:
        if not torch.equal(x, first_x): print(rep, 'INPUT x differs!', (x != first_x).nonzero()[:4].tolist())
        if not torch.equal(ref_terms, first_cpu): print(rep, 'CPU oracle differs', (ref_terms != first_cpu).nonzero()[:6].tolist())
        if not torch.equal(tg, first_gpu): print(rep, 'GPU differs', (tg != first_gpu).nonzero()[:6].tolist())
print('done')
