"""Stand-ins for torch_robotics.torch_utils.torch_utils symbols imported by
mp_baselines/planners/chomp.py:5 and cost_functions.py:14."""
import torch


def batched_weighted_dot_prod(x, M, y, with_einsum=False):
    """x, y [P,H,D], M [H,H] -> [P,D]: x[:, :, j]^T M y[:, :, j] for every particle / dim."""
    return (x.transpose(-1, -2) @ M.unsqueeze(0) @ y).diagonal(dim1=-2, dim2=-1)


def tensor_linspace_v1(start, end, steps=10):
    w = torch.linspace(0, 1, steps, dtype=start.dtype, device=start.device)
    return start.unsqueeze(-1) * (1 - w) + end.unsqueeze(-1) * w
