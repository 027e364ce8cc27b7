"""GPMP2 with the reference's constructor, attributes and return values (mp_baselines/planners/gpmp2.py).

    _step   gpmp2.py:308-342   ->  mpb_gpmp2_linearize (collision errors + analytic Jacobians, batch mean of the
                                   diagonal for the trust region) + mpb_gpmp2_solve (block-tridiagonal normal
                                   equations, block Cholesky, update, b^T K b)

The reference materialises dense A [B,rows,N], K [B,rows,rows], J^T J [B,N,N] and runs a dense Cholesky per
trajectory; the fused path never forms them (``cost.get_linear_system`` still returns the dense triplet for
callers that want it)."""
import ctypes as C

import torch

from .. import _lib
from ..costs import build_gpmp2_cost_composite
from ..factors import GPFactor, MultiMPPrior, UnaryFactor
from .base import OptimizationPlanner


class GPMP2(OptimizationPlanner):

    def __init__(self, robot=None, n_dof=None, n_support_points=None, n_interpolated_points=None,
                 num_particles_per_goal=None, opt_iters=None, dt=None, start_state=None, step_size=1.,
                 multi_goal_states=None, initial_particle_means=None, sigma_start_init=None, sigma_start_sample=None,
                 sigma_goal_init=None, sigma_goal_sample=None, sigma_gp_init=None, solver_params=None,
                 stop_criteria=None, batch_split=None, **kwargs):
        super().__init__(name='GPMP', n_dof=n_dof, n_support_points=n_support_points,
                         num_particles_per_goal=num_particles_per_goal, opt_iters=opt_iters, dt=dt,
                         start_state=start_state, initial_particle_means=initial_particle_means,
                         multi_goal_states=multi_goal_states, sigma_start_init=sigma_start_init,
                         sigma_goal_init=sigma_goal_init, sigma_gp_init=sigma_gp_init, pos_only=False, **kwargs)
        self.n_interpolated_points = n_interpolated_points
        self.robot = robot
        self.d_state_opt = 2 * self.n_dof
        self.goal_directed = multi_goal_states is not None
        if self.goal_directed and self.num_goals != 1:
            raise NotImplementedError('GPMP2 is single-goal (the reference cost breaks for > 1 goal, quirk B3)')
        self.step_size = step_size
        self.sigma_start_sample = sigma_start_sample
        self.sigma_goal_sample = sigma_goal_sample
        self.solver_params = dict(solver_params or {})
        method = self.solver_params.get('method', 'cholesky')
        if method not in ('cholesky', 'inverse', 'lstsq'):
            raise NotImplementedError(f"solver method '{method}'")
        # 'inverse' / 'lstsq' solve the same SPD system (gpmp2.py:432-491): all map to the block Cholesky
        if self.solver_params.get('sparse_computation') or self.solver_params.get('sparse_computation_block_diag'):
            raise NotImplementedError('the cholespy / sparse variants are slow fallbacks of the same maths (out of scope)')
        self.N = self.d_state_opt * self.n_support_points
        self._mean = None
        self._weights = None
        self._dist = None
        self.stop_criteria = stop_criteria
        self.costs = None
        self.cost = build_gpmp2_cost_composite(
            robot=robot, n_support_points=n_support_points, dt=dt, start_state=start_state.to(**self.tensor_args),
            multi_goal_states=None if multi_goal_states is None else multi_goal_states.to(**self.tensor_args),
            num_particles_per_goal=num_particles_per_goal, **kwargs)
        self._ws = None
        # batch_split (update.SampleSplit): the particles of ONE batch are sharded over the ranks of a process group.
        # Everything is independent per trajectory except the trust-region term, which uses the mean over the WHOLE
        # batch of diag(A^T K A) (gpmp2.py:366, quirk B10): an [H*d] all-reduce per step keeps the result identical.
        self.batch_split = batch_split
        self.reset(initial_particle_means=initial_particle_means)

    def set_prior_factors(self):
        D, ta = self.d_state_opt, self.tensor_args
        self.start_prior_init = UnaryFactor(D, self.sigma_start_init, self.start_state, ta)
        self.gp_prior_init = GPFactor(self.n_dof, self.sigma_gp_init, self.dt, self.n_support_points - 1, ta)
        if self.goal_directed:
            self.multi_goal_prior_init = [UnaryFactor(D, self.sigma_goal_init, g, ta) for g in self.multi_goal_states]
        self.start_prior_sample = UnaryFactor(D, self.sigma_start_sample, self.start_state, ta)
        if self.goal_directed:
            self.multi_goal_prior_sample = [UnaryFactor(D, self.sigma_goal_sample, g, ta) for g in self.multi_goal_states]

    def get_dist(self, start_K, gp_K, goal_K, state_init, particle_means=None, goal_states=None):
        return MultiMPPrior(self.n_support_points - 1, self.dt, 2 * self.n_dof, self.n_dof, start_K, gp_K, state_init,
                            K_g_inv=goal_K, means=particle_means, goal_states=goal_states, tensor_args=self.tensor_args)

    def reset(self, start_state=None, multi_goal_states=None, initial_particle_means=None):
        if start_state is not None:
            self.start_state = start_state.clone().to(**self.tensor_args)
        if multi_goal_states is not None:
            self.multi_goal_states = multi_goal_states.clone().to(**self.tensor_args)
        self.set_prior_factors()
        if initial_particle_means is not None:
            means = initial_particle_means.to(**self.tensor_args)
        else:
            init = self.get_dist(self.start_prior_init.K, self.gp_prior_init.Q_inv[0],
                                 self.multi_goal_prior_init[0].K if self.goal_directed else None,
                                 self.start_state, goal_states=self.multi_goal_states)
            means = init.sample(self.num_particles_per_goal)
            del init
        self._particle_means = (means.flatten(0, 1) if means.ndim == 4 else means).contiguous().clone()

    # ------------------------------------------------------------------ hot path
    def _buffers(self, B):
        H, D, d = self.n_support_points, self.d_state_opt, self.n_dof
        nf = max(1, len([c for c in self.cost.cost_l if getattr(c, 'field', None) is not None]))
        if self._ws is None or self._ws['B'] != B:
            dev = self.tensor_args['device']
            nbytes = _lib.lib().mpb_gpmp2_workspace_bytes(B, H, D)
            self._ws = dict(B=B, err=torch.zeros(nf, B, H, **self.tensor_args), hobs=torch.zeros(nf, B, H, d, **self.tensor_args),
                            dm=torch.zeros(H * d, device=dev, dtype=torch.float64),
                            ws=torch.empty(nbytes // 8, device=dev, dtype=torch.float64),
                            cost=torch.empty(B, **self.tensor_args), dtheta=torch.empty(B, H, D, **self.tensor_args))
        return self._ws

    def _step(self, debug=False, **observation):
        """One LM step in place on self._particle_means; returns the per-particle cost b^T K b at the
        linearisation point (what the reference computes from the (b, K) it returns)."""
        if observation.get('obstacle_spheres') is not None:
            raise NotImplementedError('per-call obstacle_spheres are not supported')
        B, H, d = self._particle_means.shape[0], self.n_support_points, self.n_dof
        w = self._buffers(B)
        gp, fields, nf, _ = self.cost._build()
        lib, st = _lib.lib(), _lib.stream_ptr()
        trust = bool(self.solver_params.get('trust_region', False))
        n_int, w_int = self.cost.interpolation_weights(self.n_interpolated_points)
        _lib.check(lib.mpb_gpmp2_linearize_ex(_lib.ptr(self._particle_means), B, H, C.byref(self.robot.desc), fields, nf,
                                              _lib.ptr(w['err']), _lib.ptr(w['hobs']), _lib.ptr(w['dm']) if trust else None,
                                              n_int, w_int, st))
        if trust and self.batch_split is not None and self.batch_split.world > 1:
            import torch.distributed as dist
            # local mean -> global mean: sum of (local mean * local count) over ranks / global count
            cnt = torch.tensor([float(B)], device=w['dm'].device, dtype=torch.float64)
            w['dm'].mul_(float(B))
            dist.all_reduce(w['dm'], group=self.batch_split.group)
            dist.all_reduce(cnt, group=self.batch_split.group)
            w['dm'].div_(cnt)
        inv_s2 = (C.c_float * max(1, nf))(*[fields[i].inv_sigma2 for i in range(nf)])
        _lib.check(lib.mpb_gpmp2_solve(_lib.ptr(self._particle_means), B, H, d, C.byref(gp), _lib.ptr(w['err']), _lib.ptr(w['hobs']),
                                       inv_s2, nf, _lib.ptr(w['dm']) if trust else None, float(self.solver_params['delta']),
                                       float(self.step_size), _lib.ptr(w['ws']), _lib.ptr(w['cost']), _lib.ptr(w['dtheta']), st))
        return w['cost']

    def optimize(self, opt_iters=None, debug=False, **observation):
        if opt_iters is None:
            opt_iters = self.opt_iters
        costs = costs_previous = None
        for opt_step in range(opt_iters):
            costs = self._step(debug=debug, **observation)
            if self.stop_criteria is not None:          # the only host synchronisation, as in gpmp2.py:285-293
                if opt_step == 0:
                    costs_previous = costs.clone()
                    continue
                if bool(torch.all(torch.abs((costs - costs_previous) / costs) < self.stop_criteria)):
                    break
                costs_previous = costs.clone()
        self.costs = costs.clone()
        self._recent_state_trajectories = self._particle_means[..., :self.n_dof].clone()
        self._recent_control_particles = self._particle_means[..., -self.n_dof:].clone()
        return self._get_traj()

    def _get_costs(self, errors, w_mat):
        """b^T K b for a dense (b, K) pair (gpmp2.py:493-495); kept for API parity."""
        return (errors.transpose(1, 2) @ w_mat @ errors).reshape(-1)

    def get_recent_samples(self):
        vel = self._recent_control_particles.detach().clone()
        pos = self._recent_state_trajectories.detach().clone()
        return (pos.reshape(self.num_goals, -1, *pos.shape[1:]), vel.reshape(self.num_goals, -1, *vel.shape[1:]))
