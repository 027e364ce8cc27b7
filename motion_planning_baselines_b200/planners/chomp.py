"""CHOMP with the reference's constructor, attributes and return values
(mp_baselines/planners/chomp.py); the autograd loop of ``_run_optimization`` (chomp.py:127-151) is ONE
launch of mpb_chomp_run for all ``opt_iters`` iterations (analytic collision gradient, trajectories resident
in shared memory)."""
import ctypes as C

import torch

from .. import _lib
from ..costs import CostComposite
from .base import OptimizationPlanner


class CHOMP(OptimizationPlanner):

    def __init__(self, n_dof=None, n_support_points=None, num_particles_per_goal=None, opt_iters=None, dt=None,
                 start_state=None, cost=None, weight_prior_cost=0.1, initial_particle_means=None, step_size=1.,
                 grad_clip=.01, multi_goal_states=None, sigma_start_init=0.001, sigma_goal_init=0.001,
                 sigma_gp_init=10., pos_only=False, num_particles_global=None, **kwargs):
        super().__init__(name='CHOMP', n_dof=n_dof, n_support_points=n_support_points,
                         num_particles_per_goal=num_particles_per_goal, opt_iters=opt_iters, dt=dt,
                         start_state=start_state, cost=cost, initial_particle_means=initial_particle_means,
                         multi_goal_states=multi_goal_states, sigma_start_init=sigma_start_init,
                         sigma_goal_init=sigma_goal_init, sigma_gp_init=sigma_gp_init, pos_only=pos_only, **kwargs)
        if cost is not None and not isinstance(cost, CostComposite):
            raise _lib.MpbError('CHOMP needs a motion_planning_baselines_b200.costs.CostComposite (or cost=None)')
        self.lr = step_size
        self.grad_clip = grad_clip
        self._particle_means = None
        self.Sigma_inv = self._get_R_mat(dt=self.dt, n_support_points=self.n_support_points, tensor_args=self.tensor_args)
        self.Sigma = torch.inverse(self.Sigma_inv.cpu()).to(**self.tensor_args)
        self.reset(initial_particle_means=initial_particle_means)
        self.weight_prior_cost = weight_prior_cost
        # the reference adds the smoothness cost of ALL particles to every particle (quirk B1): when the particles
        # of one problem are sharded over GPUs the global count must be given
        self.num_particles_global = num_particles_global

    @classmethod
    def _get_R_mat(cls, dt=0.01, n_support_points=64, tensor_args=None, **kwargs):
        """Backward-difference precision R = K^T K (chomp.py:81-101), built in fp32 on the host like the reference."""
        H = n_support_points
        K = torch.eye(H) - torch.diag(torch.ones(H - 1), diagonal=-1)
        K = torch.cat((K, torch.zeros(1, H)), dim=0)
        K[-1, -1] = -1.
        K = K * 1. / dt ** 2
        return (K.t() @ K).to(**tensor_args).contiguous()

    def reset(self, initial_particle_means=None):
        if initial_particle_means is not None:
            self._particle_means = initial_particle_means.to(**self.tensor_args).contiguous().clone()
        else:
            self._particle_means = self.get_random_trajs().contiguous()

    def optimize(self, opt_iters=None, **observation):
        self._run_optimization(opt_iters, **observation)
        return self._get_traj()

    def _run_optimization(self, opt_iters, **observation):
        if opt_iters is None:
            opt_iters = self.opt_iters
        if observation.get('obstacle_spheres') is not None:
            raise NotImplementedError('per-call obstacle_spheres are not supported')
        P, H = self._particle_means.shape[0], self.n_support_points
        if self.cost is None:
            raise _lib.MpbError('CHOMP needs a cost object (the robot description travels with it)')
        gp, fields, nf, _ = self.cost._build()
        if gp.enabled:
            raise NotImplementedError('the fused CHOMP gradient covers CostCollision terms (GP terms: use GPMP2)')
        P_glob = self.num_particles_global if self.num_particles_global is not None else P
        extra = self.cost._extra
        if extra is not None and extra.gp_traj_enabled:
            raise NotImplementedError('the fused CHOMP gradient covers CostCollision and CostJointLimits terms')
        if extra is not None:
            # CostJointLimits is one scalar for the whole batch that the composite adds to EVERY particle's cost; the
            # reference back-propagates costs.sum() (chomp.py:139), so its gradient carries a factor P -- the same
            # mechanism as the P-scaled smoothness term (quirk B1)
            scaled = _lib.ExtraCostDesc()
            C.memmove(C.byref(scaled), C.byref(extra), C.sizeof(scaled))
            scaled.w_jl = float(extra.w_jl) * float(P_glob)
            extra = scaled
        _lib.check(_lib.lib().mpb_chomp_run_ex(_lib.ptr(self._particle_means), P, H, C.byref(self.cost.robot.desc), fields, nf,
                                               _lib.ptr(self.Sigma_inv), float(P_glob) * float(self.weight_prior_cost),
                                               float(self.lr), float(self.grad_clip), int(opt_iters),
                                               C.byref(extra) if extra is not None else None, _lib.stream_ptr()))

    def _eval(self, x, **observation):
        """costs [P] = cost(x) + weight_prior_cost * (smoothness summed over ALL particles)  (chomp.py:153-169).
        Reporting helper (the optimisation itself never needs the value): collision part through mpb_cost_eval."""
        if x.ndim == 2:
            x = x.unsqueeze(0)
        costs = self._get_costs(x.contiguous(), **observation)
        smooth = torch.einsum('phd,hk,pkd->', x, self.Sigma_inv, x)
        return costs + self.weight_prior_cost * smooth
