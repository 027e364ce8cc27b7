import sys, numpy as np, torch
sys.path.insert(0, '.')
from motion_planning_baselines_b200 import configs
from motion_planning_baselines_b200.costs import build_gpmp2_cost_composite
from motion_planning_baselines_b200.fields import CollisionField
from motion_planning_baselines_b200.robots import Robot
dev = dict(device=torch.device('cuda:0'), dtype=torch.float32)
for cfg_name, P, S, H in [('C1', 5, 16, 64), ('C3', 4, 16, 64), ('C4', 3, 8, 64)]:
    cfg = configs.config(cfg_name)
    model, obst = cfg['robot'], cfg['obstacles']
    d = model.q_dim
    gen = torch.Generator().manual_seed(1234 + H)
    start, goal = torch.tensor(cfg['start']), torch.tensor(cfg['goal'])
    w = torch.linspace(0, 1, H).view(1, H, 1)
    line = start * (1 - w) + goal * w
    x = torch.zeros(P * S, H, 2 * d)
    x[..., :d] = line + 0.15 * torch.randn(P * S, H, d, generator=gen).cumsum(1) / np.sqrt(H) + 0.05 * torch.randn(P * S, 1, d, generator=gen)
    x[..., d:] = 0.5 * torch.randn(P * S, H, d, generator=gen)
    sig = dict(sigma_start=1e-2, sigma_gp=1.0, sigma_goal_prior=1e-2, sigma_coll=1e-1)
    robot = Robot(model, dt=cfg['dt'], tensor_args=dev)
    comp = build_gpmp2_cost_composite(robot=robot, n_support_points=H, dt=cfg['dt'], start_state=start.to(**dev),
                                      multi_goal_states=goal.to(**dev).unsqueeze(0), num_particles_per_goal=P,
                                      collision_fields=[CollisionField(obst, tensor_args=dev)], num_samples=S, tensor_args=dev, **sig)
    xg = x.to(**dev)
    ref = None
    bad = 0
    for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 300):
        terms, _ = comp.eval(xg, return_invidual_costs_and_weights=True)
        t = torch.stack(terms).cpu()
        if ref is None:
            ref = t
        elif not torch.equal(t, ref):
            bad += 1
            diff = (t != ref).nonzero()
            if bad <= 3:
                print(cfg_name, 'iter', it, 'differs at (term,sample):', diff.tolist()[:6], 'vals', [(float(t[i, j]), float(ref[i, j])) for i, j in diff.tolist()[:3]])
    print(cfg_name, 'runs differing from first:', bad, 'of 299')
