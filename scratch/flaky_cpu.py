import sys, numpy as np, torch
sys.path.insert(0, '.')
from motion_planning_baselines_b200 import configs
from oracle.build import TA, oracle_field, oracle_robot
from oracle.costs import CostSpec
cfg_name, P, S, H = 'C1', 5, 16, 64
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
print('threads', torch.get_num_threads(), 'x checksum', float(x.double().sum()))
spec = CostSpec(oracle_robot(model, cfg['dt']), H, cfg['dt'], start, goal, [oracle_field(obst, model)], tensor_args=TA, **sig)
ref = None
bad = 0
for it in range(int(sys.argv[1])):
    t = torch.stack(spec.terms(x))
    if ref is None:
        ref = t
        print('first', [float(t[2, i]) for i in (30, 31, 33, 34)])
    elif not torch.equal(t, ref):
        bad += 1
        if bad < 4:
            dd = (t != ref).nonzero().tolist()
            print('iter', it, 'diff at', dd[:6], [(float(t[i, j]), float(ref[i, j])) for i, j in dd[:3]])
print('differing runs', bad)
