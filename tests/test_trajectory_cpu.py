"""CPU: the batched trajectory helpers behind HybridPlanner (motion_planning_baselines_b200/trajectory.py; the
reference imports them from the absent torch_robotics -- hybrid_planner.py:5-7 -- so this file pins OUR specification):
piecewise-linear re-sampling uniform in the waypoint index, average-velocity fill, central differences."""
import numpy as np
import torch

from motion_planning_baselines_b200 import trajectory as tr


def test_resample_paths_matches_numpy_interp():
    gen = torch.Generator().manual_seed(0)
    paths = [torch.randn(n, 3, generator=gen) for n in (2, 5, 17, 1, 64)]
    H = 33
    out = tr.resample_paths(paths, H)
    assert out.shape == (5, H, 3)
    for i, p in enumerate(paths):
        n = p.shape[0]
        if n == 1:
            want = p.numpy().repeat(H, 0)
        else:
            want = np.stack([np.interp(np.linspace(0, n - 1, H), np.arange(n), p[:, k].numpy()) for k in range(3)], axis=1)
        assert np.abs(out[i].numpy() - want).max() < 1e-6
        assert torch.equal(out[i, 0], p[0]) and torch.equal(out[i, -1], p[-1])


def test_smoothen_trajectory_velocity_conventions():
    p = torch.tensor([[0., 0.], [1., 0.], [1., 2.]])
    pos, vel = tr.smoothen_trajectory(p, n_support_points=9, dt=0.5)
    assert pos.shape == vel.shape == (9, 2)
    assert torch.allclose(vel[1:-1], torch.tensor([1., 2.]).expand(7, 2) / (9 * 0.5)) and float(vel[[0, -1]].abs().max()) == 0
    _, v0 = tr.smoothen_trajectory(p, n_support_points=9, dt=0.5, zero_velocity=True)
    assert float(v0.abs().max()) == 0
    pos2, vfd = tr.smoothen_trajectory(p, n_support_points=9, dt=0.5, set_average_velocity=False)
    assert torch.allclose(vfd[1:-1], (pos2[2:] - pos2[:-2]) / 1.0)


def test_tensor_linspace_layout():
    a, b = torch.tensor([0., 1.]), torch.tensor([2., 5.])
    ls = tr.tensor_linspace(a, b, 5)
    assert ls.shape == (2, 5)                      # the reference transposes it into [steps, d]
    assert torch.allclose(ls.T[0], a) and torch.allclose(ls.T[-1], b) and torch.allclose(ls.T[2], (a + b) / 2)
