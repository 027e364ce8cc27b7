"""CPU: hand-checkable known answers for the oracle's collision fields (oracle/fields.py).  These fields are the
specification at the torch_robotics boundary (parity unpinned, see oracle/__init__.py), so what can be pinned is
pinned by construction: closed-form distances, hinge values and sub-gradients."""
import math

import numpy as np
import torch

from oracle.fields import PrimitiveField, SelfCollisionField, WorkspaceBoundaryField


def test_primitive_field_known_answers():
    f = PrimitiveField(sphere_centers=[[0., 0., 0.]], sphere_radii=[0.5], box_centers=[[2., 0., 0.]], box_half=[[0.5, 1., 1.]],
                       link_radii=[0.1], cutoff_margin=0.05)
    x = torch.tensor([[1.0, 0., 0.],      # 0.5 from the sphere, 0.5 from the box face
                      [0.0, 0., 0.],      # sphere centre: -0.5
                      [2.0, 0., 0.],      # box centre: -0.5
                      [1.2, 0., 0.],      # 0.3 from the box, 0.7 from the sphere
                      [1.0, 2., 2.]])     # outside the box corner region: sqrt(0.25+1+1)=1.5 from the box
    sdf = f.sdf(x)
    assert torch.allclose(sdf, torch.tensor([0.5, -0.5, -0.5, 0.3, 1.5]), atol=1e-6)
    h = f.hinge_terms(x.unsqueeze(-2)).squeeze(-1)
    assert torch.allclose(h, torch.tensor([0., 0.65, 0.65, 0., 0.]), atol=1e-6)


def test_self_collision_field_known_answers():
    # three spheres on a line at x = 0, 0.3, 1.0 with radii 0.1, 0.1, 0.2; pairs (0,1), (0,2), (1,2); margin 0.05
    f = SelfCollisionField([[0, 1], [0, 2], [1, 2]], [0.1, 0.1, 0.2], cutoff_margin=0.05)
    c = torch.tensor([[[0., 0., 0.], [0.3, 0., 0.], [1.0, 0., 0.]]], requires_grad=True)
    h = f.hinge_terms(c)
    assert torch.allclose(h, torch.tensor([[0., 0., 0.]]))                    # 0.25-0.3<0, 0.35-1<0, 0.35-0.7<0
    c2 = torch.tensor([[[0., 0., 0.], [0.2, 0., 0.], [0.4, 0., 0.]]], requires_grad=True)
    h2 = f.hinge_terms(c2)
    assert torch.allclose(h2, torch.tensor([[0.05, 0., 0.15]]), atol=1e-6)    # 0.25-0.2, 0.35-0.4<0, 0.35-0.2
    cost = f.compute_cost(None, c2)
    assert math.isclose(float(cost.detach()), 0.2, rel_tol=1e-5)
    g, = torch.autograd.grad(cost.sum(), c2)
    # moving sphere 0 towards +x (closer to sphere 1) raises the cost; sphere 1 sits between two active pairs; sphere 2
    # lowers the cost by moving to +x
    assert torch.allclose(g[0, :, 0], torch.tensor([1., 0., -1.]))
    assert float(g[0, :, 1:].abs().max()) == 0.0
    assert not bool(f.collision_free(c2.detach().unsqueeze(0))[0]) and bool(f.collision_free(c.detach().unsqueeze(0))[0])


def test_workspace_field_known_answers():
    f = WorkspaceBoundaryField([-1., -1., 0.], [1., 1., 2.], link_radii=[0.1], cutoff_margin=0.05)
    x = torch.tensor([[0., 0., 1.],        # 1 from every wall
                      [0.9, 0., 1.],       # 0.1 from +x wall
                      [0., -1.2, 1.],      # 0.2 outside the -y wall
                      [0.5, 0.5, 0.5]], requires_grad=True)   # tie between +x, +y (0.5) and z-lo (0.5): first entry wins
    sdf = f.sdf(x)
    assert torch.allclose(sdf, torch.tensor([1., 0.1, -0.2, 0.5]), atol=1e-6)
    h = f.hinge_terms(x.unsqueeze(-2)).squeeze(-1)
    assert torch.allclose(h, torch.tensor([0., 0.05, 0.35, 0.]), atol=1e-6)
    g, = torch.autograd.grad(sdf.sum(), x)
    assert torch.equal(g[1], torch.tensor([-1., 0., 0.])) and torch.equal(g[2], torch.tensor([0., 1., 0.]))
    assert torch.equal(g[3], torch.tensor([0., 0., 1.])), 'first minimal entry in the order (x-lo, hi-x) is z - lo_z'
    f2 = WorkspaceBoundaryField([-1., -1.], [1., 1.], link_radii=[0.0], cutoff_margin=0.1)
    assert torch.allclose(f2.compute_cost(None, torch.tensor([[[[0.95, 0.]]]])), torch.tensor([[0.05]]), atol=1e-6)


def test_panda_self_collision_pair_table():
    from motion_planning_baselines_b200.models import PANDA_SELF_LINK_PAIRS, panda_model, self_collision_pairs
    m = panda_model()
    pairs = self_collision_pairs(m)
    link = np.asarray(m.sphere_link)
    assert pairs.shape == (564, 2)
    lp = sorted(set((int(link[i]), int(link[j])) for i, j in pairs))
    assert lp == sorted(PANDA_SELF_LINK_PAIRS)
    keys = link[pairs[:, 0]] * 8 + link[pairs[:, 1]]
    assert (np.diff(keys) >= 0).all(), 'pairs must be sorted by link pair (the C ABI contract)'
    assert all(b - a >= 2 for a, b in lp)
