"""Oracle GP trajectory prior (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Restates, as plain functions, the maths of
  * GPFactor.calc_phi / calc_Q_inv     mp_baselines/planners/costs/factors/gp_factor.py:34-50
  * UnaryFactor.K                      mp_baselines/planners/costs/factors/unary_factor.py:19
  * MultiMPPrior.get_const_vel_covariance / const_vel_trajectory / update_dist / sample
                                       mp_baselines/planners/costs/factors/mp_priors_multi.py:100-110,130-151,213-256
  * torch.distributions.MultivariateNormal(precision_matrix=...) -> scale_tril, rsample
Pinned by tests/golden/prior_*.npz (generated from the unmodified reference).
"""
import torch


def phi_matrix(d, dt, tensor_args):
    """[[I, dt I], [0, I]]  (gp_factor.py:34-40)."""
    Phi = torch.eye(2 * d, **tensor_args)
    Phi[:d, d:] = torch.eye(d, **tensor_args) * dt
    return Phi


def gp_Q_inv(d, dt, sigma_gp, tensor_args):
    """[[12 dt^-3, -6 dt^-2], [-6 dt^-2, 4 dt^-1]] (x) I/sigma^2 in the working dtype
    (gp_factor.py:23-26,42-50)."""
    Qc_inv = torch.eye(d, **tensor_args) / sigma_gp ** 2
    a = 12. * (dt ** -3.) * Qc_inv
    b = -6. * (dt ** -2.) * Qc_inv
    c = 4. * (dt ** -1.) * Qc_inv
    return torch.cat((torch.cat((a, b), dim=-1), torch.cat((b, c), dim=-1)), dim=-2)


def unary_K(D, sigma, tensor_args):
    """I / sigma^2 (unary_factor.py:19)."""
    return torch.eye(D, **tensor_args) / sigma ** 2


def prior_precision(H, d, dt, K_s, Q_inv, K_g, tensor_args):
    """Sigma^-1 = A^T Qfull^-1 A assembled in fp64, cast to the working dtype
    (mp_priors_multi.py:213-251).  K_s / Q_inv / K_g arrive in the working dtype
    (already rounded) exactly like in the reference."""
    D = 2 * d
    M = D * H
    f64 = dict(device=tensor_args['device'], dtype=torch.float64)
    Phi = phi_matrix(d, dt, f64)
    A = torch.eye(M, **f64)
    A[D:, :-D] -= torch.kron(torch.eye(H - 1, **f64), Phi)
    blocks = [K_s.to(**f64)] + [Q_inv.to(**f64)] * (H - 1)
    if K_g is not None:
        goal_rows = torch.zeros(D, M, **f64)
        goal_rows[:, -D:] = torch.eye(D, **f64)
        A = torch.cat((A, goal_rows))
        blocks.append(K_g.to(**f64))
    Qfull = torch.block_diag(*blocks)
    return (A.t() @ Qfull @ A).to(**tensor_args)


def precision_to_scale_tril(P):
    """What MultivariateNormal(precision_matrix=P) stores as ``scale_tril``
    (torch/distributions/multivariate_normal.py ``_precision_to_scale_tril``)."""
    Lf = torch.linalg.cholesky(torch.flip(P, (-2, -1)))
    L_inv = torch.transpose(torch.flip(Lf, (-2, -1)), -2, -1)
    eye = torch.eye(P.shape[-1], dtype=P.dtype, device=P.device)
    return torch.linalg.solve_triangular(L_inv, eye, upper=False)


def const_vel_mean(start_state, goal_state, dt, H, d, tensor_args, zero_end_vel=True):
    """Straight line start->goal, constant velocity on interior waypoints and zero
    velocity at both ends (mp_priors_multi.py:130-151)."""
    n = H - 1
    traj = torch.zeros(H, 2 * d, **tensor_args)
    mean_vel = (goal_state[:d] - start_state[:d]) / (n * dt)
    for i in range(H):
        traj[i, :d] = start_state[:d] * (n - i) * 1. / n + goal_state[:d] * i * 1. / n
    if zero_end_vel:
        traj[1:-1, d:] = mean_vel
    else:
        traj[:, d:] = mean_vel
    return traj


def sample_prior(means, L, eps):
    """x[p,s] = means[p] + L @ eps[s,p]  (MultivariateNormal.rsample + the view/transpose of
    mp_priors_multi.py:253-256).  means [P,M], L [M,M], eps [S,P,M] -> [P,S,M]."""
    x = means.unsqueeze(0) + eps @ L.transpose(-1, -2)
    return x.transpose(0, 1)


def sample_prior_faithful(means, L_batched, eps):
    """Same numbers, the reference's cost model: a broadcast batched mat-vec against a
    per-particle copy of the factor (torch ``_batch_mv`` with scale_tril [P,M,M])."""
    x = means.unsqueeze(0) + torch.matmul(L_batched, eps.unsqueeze(-1)).squeeze(-1)
    return x.transpose(0, 1)
