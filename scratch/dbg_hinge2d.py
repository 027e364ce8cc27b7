import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_num_threads(1)
from motion_planning_baselines_b200 import configs
from motion_planning_baselines_b200.costs import CostCollision, CostComposite
from motion_planning_baselines_b200.fields import CollisionField
from motion_planning_baselines_b200.robots import Robot
from oracle.build import TA, oracle_field
dev = dict(device=torch.device('cuda:0'), dtype=torch.float32)
for name in ('C1', 'C2', 'C3'):
    cfg = configs.config(name)
    model = cfg['robot']; d = model.q_dim
    robot = Robot(model, dt=0.1, tensor_args=dev)
    f = CollisionField(cfg['obstacles'], tensor_args=dev)
    of = oracle_field(cfg['obstacles'], model)
    gen = torch.Generator().manual_seed(1)
    B = 20000
    x = torch.zeros(B, 2, 2 * d)
    x[:, 1, :d] = 2.2 * torch.rand(B, d, generator=gen) - 1.1
    cost = CostComposite(robot, 2, [CostCollision(robot, 2, field=f, sigma_coll=1.0, tensor_args=dev)], tensor_args=dev)
    got = cost.eval(x.to(**dev)).cpu()
    q = x[:, 1:, :d]
    ref = of.compute_cost(q, q.unsqueeze(-2)).reshape(B)
    bad = (got != ref).nonzero().flatten()
    print(name, 'mismatches', len(bad), 'of', B, 'nonzero', int((ref > 0).sum()))
    for i in bad[:8].tolist():
        p = q[i, 0]
        sph = None
        if of.sphere_centers.shape[0]:
            ds = (p - of.sphere_centers).norm(dim=-1) - of.sphere_radii
            sph = (int(ds.argmin()), float(ds.min()))
        box = None
        if of.box_centers.shape[0]:
            qq = (p - of.box_centers).abs() - of.box_half
            box = (int(qq.max(-1).values.argmin()), qq[qq.max(-1).values.argmin()].tolist())
        print('  ', i, p.tolist(), 'got %.9g ref %.9g' % (got[i], ref[i]), 'sphere', sph, 'box', box)
