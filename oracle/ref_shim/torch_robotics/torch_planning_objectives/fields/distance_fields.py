"""Stand-in for interpolate_points_v1 (mp_baselines/planners/costs/cost_functions.py:13,118):
linear joint-space up-sampling with n extra points between consecutive waypoints."""
import torch


def interpolate_points_v1(trajs, n_interpolated_points):
    B, H, D = trajs.shape
    w = torch.linspace(0, 1, n_interpolated_points + 2, dtype=trajs.dtype, device=trajs.device)[:-1]
    seg = trajs[:, :-1].unsqueeze(2) * (1 - w).view(1, 1, -1, 1) + trajs[:, 1:].unsqueeze(2) * w.view(1, 1, -1, 1)
    return torch.cat((seg.reshape(B, -1, D), trajs[:, -1:]), dim=1)
