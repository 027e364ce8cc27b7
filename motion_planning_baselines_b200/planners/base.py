"""Planner shells with the reference's public surface (mp_baselines/planners/base.py):
``optimize(opt_iters, **observation)``, ``__call__``, ``get_traj``, particle bookkeeping."""
import torch

from .. import _lib
from ..factors import GPFactor, MultiMPPrior, UnaryFactor


class MPPlanner:
    def __init__(self, name=None, tensor_args=None, **kwargs):
        self.name = name
        if tensor_args is None or torch.device(tensor_args['device']).type != 'cuda':
            raise _lib.MpbError("tensor_args['device'] must be a CUDA device: the hot path has no CPU implementation")
        if tensor_args.get('dtype', torch.float32) != torch.float32:
            raise _lib.MpbError('the fused hot path computes in float32')
        self.tensor_args = dict(device=torch.device(tensor_args['device']), dtype=torch.float32)
        _lib.init_device(self.tensor_args['device'])
        self._kwargs = kwargs

    def optimize(self, opt_iters=1, **observation):
        raise NotImplementedError

    def __call__(self, opt_iters=1, **observation):
        return self.optimize(opt_iters, **observation)

    def __repr__(self):
        return f'{self.name}({self._kwargs})'

    def render(self, ax, **kwargs):
        raise NotImplementedError


class OptimizationPlanner(MPPlanner):
    """Dimensions / particle counts / zero-velocity concatenation of base.py:60-113."""

    def __init__(self, name='OptimizationPlanner', n_dof=None, n_support_points=None, n_interpolated_points=None,
                 num_particles_per_goal=None, opt_iters=None, dt=None, start_state=None, cost=None,
                 initial_particle_means=None, multi_goal_states=None, sigma_start_init=0.001, sigma_goal_init=0.001,
                 sigma_gp_init=10., pos_only=False, tensor_args=None, **kwargs):
        super().__init__(name, tensor_args, **kwargs)
        self.n_dof = n_dof
        self.dim = 2 * n_dof
        self.n_support_points = n_support_points
        self.n_interpolated_points = n_interpolated_points
        self.num_particles_per_goal = num_particles_per_goal
        self.opt_iters = opt_iters
        self.dt = dt
        self.pos_only = pos_only
        if pos_only:
            raise NotImplementedError('pos_only=True is not supported (reference quirk B5 makes it unusable there too)')
        self.start_state = start_state.to(**self.tensor_args)
        self.multi_goal_states = None if multi_goal_states is None else multi_goal_states.to(**self.tensor_args)
        if multi_goal_states is None:
            self.num_goals = 1
        else:
            assert multi_goal_states.ndim == 2
            self.num_goals = multi_goal_states.shape[0]
        self.num_particles = self.num_goals * num_particles_per_goal
        self.cost = cost
        self.initial_particle_means = initial_particle_means
        self._particle_means = None
        self.d_state_opt = 2 * n_dof
        self.start_state = torch.cat([self.start_state, torch.zeros_like(self.start_state)], dim=-1)
        if self.multi_goal_states is not None:
            self.multi_goal_states = torch.cat([self.multi_goal_states, torch.zeros_like(self.multi_goal_states)], dim=-1)
        self.sigma_start_init = sigma_start_init
        self.sigma_goal_init = sigma_goal_init
        self.sigma_gp_init = sigma_gp_init

    def get_GP_prior(self, start_K, gp_K, goal_K, state_init, particle_means=None, goal_states=None, tensor_args=None):
        return MultiMPPrior(self.n_support_points - 1, self.dt, self.dim, self.n_dof, start_K, gp_K, state_init,
                            K_g_inv=goal_K, means=particle_means, goal_states=goal_states,
                            tensor_args=tensor_args or self.tensor_args)

    def get_random_trajs(self, eps=None):
        """Initial particles ~ GP prior around the straight line (base.py:155-202).  The reference forces fp64 for
        this one-off step (quirk B8): factors, precision and the sampling factor are built in fp64 here too (host side,
        once) and only rounded to fp32 for the device sampler, which keeps the particles within ~1e-6 of the
        reference's (tests/test_gpu_init_path.py).  ``eps``: optional injected noise [P_per_goal, num_goals, M]."""
        f64 = dict(device='cpu', dtype=torch.float64)
        start = torch.cat((self.start_state, torch.zeros_like(self.start_state)), dim=-1)
        goals = None
        if self.multi_goal_states is not None:
            goals = torch.cat((self.multi_goal_states, torch.zeros_like(self.multi_goal_states)), dim=-1)
        D, d, dt = 2 * self.n_dof, self.n_dof, self.dt
        self.start_prior_init = UnaryFactor(D, self.sigma_start_init, start, self.tensor_args)
        self.gp_prior_init = GPFactor(self.n_dof, self.sigma_gp_init, self.dt, self.n_support_points - 1, self.tensor_args)
        K_s = torch.eye(D, **f64) / self.sigma_start_init ** 2
        Qc = torch.eye(d, **f64) / self.sigma_gp_init ** 2
        a, b, c = 12. * (dt ** -3.) * Qc, -6. * (dt ** -2.) * Qc, 4. * (dt ** -1.) * Qc
        Q = torch.cat((torch.cat((a, b), dim=-1), torch.cat((b, c), dim=-1)), dim=-2)
        K_g = torch.eye(D, **f64) / self.sigma_goal_init ** 2 if goals is not None else None
        prior = MultiMPPrior(self.n_support_points - 1, self.dt, self.dim, self.n_dof, K_s, Q, start, K_g_inv=K_g,
                             goal_states=goals, tensor_args=self.tensor_args, factor_dtype=torch.float64)
        if eps is not None:
            eps = eps.to(**self.tensor_args).contiguous()
        particles = prior.sample(self.num_particles_per_goal, eps=eps)
        self.traj_dim = particles.shape
        return particles.flatten(0, 1).clone()

    def _get_traj(self):
        return self._particle_means.clone()

    def get_traj(self):
        return self._get_traj()

    def _get_costs(self, state_trajectories, **observation):
        if self.cost is None:
            return torch.zeros(self.num_particles, **self.tensor_args)
        return self.cost(state_trajectories, **observation)
