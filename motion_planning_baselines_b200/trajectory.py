"""Trajectory helpers the reference imports from the absent torch_robotics (``tensor_linspace_v1``,
``smoothen_trajectory``, ``finite_difference_vector``: mp_baselines/planners/hybrid_planner.py:5-7,50-57,
base.py:9,211).  PARITY UNPINNED: none of them is in /root/reference, so the behaviour below is this package's
specification: paths are re-sampled piecewise-linearly, uniformly in the waypoint index, and
``set_average_velocity`` writes (end - start) / (H dt) on the interior waypoints with zero velocity at both ends.
Everything is batched on the device: B paths of different lengths are padded into one tensor and re-sampled by one
gather + lerp instead of the reference's per-path python loop (hybrid_planner.py:46-62)."""
import torch


def tensor_linspace(start, end, steps):
    """[..., d] x [..., d] -> [..., d, steps] (the layout of torch_robotics' tensor_linspace_v1; the reference
    transposes it into [steps, d], hybrid_planner.py:50-53)."""
    w = torch.linspace(0, 1, steps, device=start.device, dtype=start.dtype)
    return start.unsqueeze(-1) * (1 - w) + end.unsqueeze(-1) * w


def finite_difference_vector(x, dt=1.0):
    """Central differences on the interior waypoints, zero at both ends ([..., H, d] -> [..., H, d])."""
    v = torch.zeros_like(x)
    v[..., 1:-1, :] = (x[..., 2:, :] - x[..., :-2, :]) / (2 * dt)
    return v


def resample_paths(paths, n_support_points, tensor_args=None):
    """List of B position paths [n_i, d] (n_i >= 1, may differ) -> [B, n_support_points, d]: piecewise-linear
    re-sampling, uniform in the waypoint index, end points kept exactly."""
    B = len(paths)
    d = paths[0].shape[-1]
    ta = tensor_args if tensor_args is not None else dict(device=paths[0].device, dtype=paths[0].dtype)
    n = torch.tensor([p.shape[0] for p in paths], device=ta['device'])
    n_max = int(n.max())
    padded = torch.zeros(B, n_max + 1, d, **ta)
    for i, p in enumerate(paths):                       # B small host-side copies; everything after is batched
        padded[i, :p.shape[0]] = p.to(**ta)
        padded[i, p.shape[0]:] = p[-1].to(**ta)
    u = torch.linspace(0, 1, n_support_points, **ta).unsqueeze(0) * (n - 1).unsqueeze(1).to(ta['dtype'])   # [B,H]
    i0 = u.floor().long().clamp_(min=0)
    i0 = torch.minimum(i0, (n - 1).clamp(min=0).unsqueeze(1))
    frac = (u - i0.to(ta['dtype'])).unsqueeze(-1)
    a = torch.gather(padded, 1, i0.unsqueeze(-1).expand(B, n_support_points, d))
    b = torch.gather(padded, 1, (i0 + 1).unsqueeze(-1).expand(B, n_support_points, d))
    out = a + (b - a) * frac
    out[:, -1] = padded[torch.arange(B, device=ta['device']), n - 1]
    return out


def smoothen_trajectory(traj_pos, n_support_points=30, dt=0.02, set_average_velocity=True, zero_velocity=False,
                        tensor_args=None):
    """One path [n, d] -> (pos [H, d], vel [H, d]).  See the module docstring for the (unpinned) semantics."""
    pos, vel = smoothen_trajectories([traj_pos], n_support_points, dt, set_average_velocity, zero_velocity, tensor_args)
    return pos[0], vel[0]


def smoothen_trajectories(paths, n_support_points=30, dt=0.02, set_average_velocity=True, zero_velocity=False,
                          tensor_args=None):
    """Batched form: list of B paths -> (pos [B,H,d], vel [B,H,d])."""
    pos = resample_paths(paths, n_support_points, tensor_args)
    vel = torch.zeros_like(pos)
    if zero_velocity:
        pass
    elif set_average_velocity:
        vel[:, 1:-1] = ((pos[:, -1] - pos[:, 0]) / (n_support_points * dt)).unsqueeze(1)
    else:
        vel = finite_difference_vector(pos, dt=dt * 1.0)
    return pos, vel
