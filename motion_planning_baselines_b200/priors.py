"""Per-control-dimension Gaussian prior over control trajectories with the reference's names
(mp_baselines/planners/priors/gaussian.py:85-198,218-333).  Covariances are built exactly like the reference does;
the Cholesky factors the sampler multiplies by come from torch's own MultivariateNormal(covariance_matrix=...)
routine, once, on the host.  Drawing samples is part of the fused MPPI kernel (csrc/mppi.cu)."""
import numpy as np
import torch
import torch.distributions as dist

_CPU32 = dict(device='cpu', dtype=torch.float32)


def diag_Cov(sigma, length=None, ctrl_dim=None, tensor_args=None):
    """Time-independent diagonal covariance [T,T,C] (gaussian.py:143-163)."""
    Cov = torch.eye(length, **_CPU32).unsqueeze(-1).repeat(1, 1, ctrl_dim)
    if isinstance(sigma, (list, tuple)):
        Cov = Cov * torch.Tensor(np.array(sigma)).to(**_CPU32) ** 2
    else:
        Cov = Cov * sigma ** 2
    return Cov.to(**tensor_args)


def const_ctrl_Cov(sigma, length=None, ctrl_dim=None, tensor_args=None):
    """Constant-control covariance sigma^2 (L1 L1^T + 1 1^T), L1 = strictly-lower ones [T,T-1] (gaussian.py:166-198)."""
    if isinstance(sigma, (list, tuple)):
        sigma = torch.from_numpy(np.array(sigma)).to(**_CPU32)
    L = torch.tril(torch.ones(length, length - 1, **_CPU32), diagonal=-1)
    LL_t = torch.matmul(L, L.transpose(0, 1))
    LL_t += torch.ones(length, length, **_CPU32)
    Cov = LL_t.unsqueeze(-1).repeat(1, 1, ctrl_dim) * sigma ** 2
    return Cov.to(**tensor_args)


class ControlTrajectoryGaussian:
    """mu [T,C], Cov [T,T,C]; ``scale_tril`` [C,T,T] are the factors of the per-dimension MultivariateNormal
    distributions the reference keeps in ``list_ctrl_dists`` (gaussian.py:301-333)."""

    def __init__(self, rollout_steps, ctrl_dim, mu=None, Cov=None, tensor_args=None):
        assert mu.size(0) == rollout_steps and mu.size(1) == ctrl_dim
        self.rollout_steps, self.ctrl_dim, self.tensor_args = rollout_steps, ctrl_dim, tensor_args
        self.mu = mu
        self.Cov = Cov
        Cov_cpu = Cov.detach().to(**_CPU32)
        self.scale_tril = torch.stack([
            dist.MultivariateNormal(torch.zeros(rollout_steps), covariance_matrix=Cov_cpu[:, :, i]).scale_tril
            for i in range(ctrl_dim)]).to(**tensor_args).contiguous()

    def update_means(self, means):
        self.mu = means.detach().clone()


def get_multivar_gaussian_prior(sigma, rollout_steps, control_dim, Cov_type='indep_ctrl', mu_init=None, tensor_args=None):
    assert Cov_type in ('indep_ctrl', 'const_ctrl'), 'Invalid type for control prior dist.'
    mu = torch.zeros(rollout_steps, control_dim, **tensor_args)
    if mu_init is not None:
        mu[:, :] = mu_init
    Cov = (const_ctrl_Cov if Cov_type == 'const_ctrl' else diag_Cov)(sigma, rollout_steps, control_dim, tensor_args=tensor_args)
    return ControlTrajectoryGaussian(rollout_steps, control_dim, mu, Cov, tensor_args=tensor_args)
