"""STOMP with the reference's constructor, attributes and return values
(mp_baselines/planners/stomp.py), its iteration body replaced by the fused kernels:

    sample                 stomp.py:97-108    ->  mpb_sample_stomp  (L_R @ eps, endpoint zeroing, + mean)
    _sample_and_eval       stomp.py:162-197   ->  mpb_sample_stomp + cost.eval (mpb_cost_eval)
    _update_distribution   stomp.py:199-220   ->  mpb_softmax_update with SigmaR = inverse(R)
"""
import ctypes as C

import torch
import torch.distributions as dist

from .. import _lib
from ..update import SampleSplit, split_softmax_update
from .base import OptimizationPlanner


class STOMP(OptimizationPlanner):
    SPLIT_THRESHOLD = 4096      # samples per particle above which the update is split over several CTAs

    def __init__(self, n_dof=None, n_support_points=None, num_particles_per_goal=None, num_samples=None,
                 opt_iters=None, dt=None, start_state=None, cost=None, initial_particle_means=None,
                 multi_goal_states=None, sigma_start_init=0.001, sigma_goal_init=0.001, sigma_gp_init=10.,
                 temperature=1., step_size=1., sigma_spectral=0.1, goal_state=None, pos_only=False,
                 tensor_args=None, sample_split=None, seed=None, noise_particle_offset=0, noise_particles_global=None,
                 **kwargs):
        super().__init__(name='STOMP', n_dof=n_dof, n_support_points=n_support_points,
                         num_particles_per_goal=num_particles_per_goal, opt_iters=opt_iters, dt=dt,
                         start_state=start_state, cost=cost, initial_particle_means=initial_particle_means,
                         multi_goal_states=multi_goal_states, sigma_start_init=sigma_start_init,
                         sigma_goal_init=sigma_goal_init, sigma_gp_init=sigma_gp_init, pos_only=pos_only,
                         tensor_args=tensor_args)
        self.lr = step_size
        self.sigma_spectral = sigma_spectral
        self.start_state = start_state.to(**self.tensor_args)        # the reference undoes the base-class concat (quirk B7)
        self.goal_state = goal_state
        self.num_samples = num_samples
        self.temperature = temperature
        self._particle_means = None
        self._weights = None
        self._sample_dist = None
        self.costs = None

        R_cpu = self._get_R_mat_cpu()
        self.Sigma_inv = R_cpu.to(**self.tensor_args).contiguous()
        self.Sigma = torch.inverse(R_cpu).to(**self.tensor_args).contiguous()
        # factor of the noise distribution from the reference's own routine (stomp.py:88-95), once, on the CPU
        self._L_R = dist.MultivariateNormal(torch.zeros(n_support_points), precision_matrix=R_cpu).scale_tril \
            .to(**self.tensor_args).contiguous()
        # sample-split mode: the S samples of every particle are sharded over the ranks of a process group
        self.split = sample_split or SampleSplit(world=1, rank=0)
        self._offset, self._s_local = self.split.local_slice(num_samples)
        # in-kernel noise: one global Philox stream [S_glob, D, P_glob, H]; a rank of a sample split draws its own samples
        P_glob = noise_particles_global if noise_particles_global is not None else noise_particle_offset + self.num_particles
        self._noise = _lib.NoiseStream(seed, p_offset=noise_particle_offset, P_global=P_glob, s_offset=self._offset)
        P, S, H, D = self.num_particles, self._s_local, n_support_points, self.d_state_opt
        self.state_particles = torch.empty(P, S, H, D, **self.tensor_args)
        self._w_buf = torch.empty(P, S, **self.tensor_args)
        self._cost_buf = torch.empty(P * S, **self.tensor_args)
        self.reset(initial_particle_means=initial_particle_means)
        self.best_cost = torch.inf

    def _get_R_mat_cpu(self):
        """Second-difference precision R = A^T A (stomp.py:68-86), built in fp32 on the host like the reference."""
        H = self.n_support_points
        A = torch.diag(torch.ones(H - 1), diagonal=1) + torch.diag(torch.ones(H - 1), diagonal=-1) - 2 * torch.eye(H)
        A = torch.cat((torch.zeros(1, H), A, torch.zeros(1, H)), dim=0)
        A[0, 0] = 1.
        A[-1, -1] = 1.
        A = A * 1. / self.dt ** 2 * self.sigma_spectral
        return A.t() @ A

    def _get_R_mat(self):
        return self.Sigma_inv

    def set_noise_dist(self):
        pass            # the factor L_R is fixed at construction; nothing to rebuild

    def sample(self, eps=None):
        """-> [P,S_local,H,D]; ``eps`` [S,D,P,H] is the block torch would draw (stomp.py:102); with a sample split
        every rank consumes its own slice of the sample axis."""
        P, S, H, D = self.num_particles, self._s_local, self.n_support_points, self.d_state_opt
        if eps is None:     # drawn inside the kernel (Philox keyed on the global element index)
            nd = self._noise.next()
            _lib.check(_lib.lib().mpb_sample_stomp_rng(_lib.ptr(self._L_R), _lib.ptr(self._particle_means), C.byref(nd),
                                                       _lib.ptr(self.state_particles), P, S, H, D, _lib.stream_ptr()))
            return self.state_particles
        _lib.require_f32(eps)
        assert eps.shape == (self.num_samples, D, P, H)
        eps = eps[self._offset:self._offset + S].contiguous()
        _lib.check(_lib.lib().mpb_sample_stomp(_lib.ptr(self._L_R), _lib.ptr(self._particle_means), _lib.ptr(eps),
                                               _lib.ptr(self.state_particles), P, S, H, D, _lib.stream_ptr()))
        return self.state_particles

    def reset(self, initial_particle_means=None, eps=None):
        if initial_particle_means is not None:
            self._particle_means = initial_particle_means.to(**self.tensor_args).contiguous().clone()
        else:
            self._particle_means = self.get_random_trajs().contiguous()
        self.state_particles = self.sample(eps=eps)

    def optimize(self, opt_iters=None, eps=None, **observation):
        self._run_optimization(opt_iters, eps=eps, **observation)
        return self._get_traj()

    def _run_optimization(self, opt_iters, eps=None, **observation):
        if opt_iters is None:
            opt_iters = self.opt_iters
        if self._can_run_fused(eps, observation):
            # all iterations enqueued from one C call (three launches each, no host work in between): the small STOMP
            # configurations are launch-latency bound
            P, S, H, D = self.num_particles, self._s_local, self.n_support_points, self.d_state_opt
            gp, fields, nf, _ = self.cost._build()
            nd = self._noise.desc()
            _lib.check(_lib.lib().mpb_stomp_run(
                _lib.ptr(self._L_R), _lib.ptr(self.Sigma), C.byref(nd), _lib.ptr(self._particle_means), _lib.ptr(self.state_particles),
                _lib.ptr(self._cost_buf), _lib.ptr(self._w_buf), P, S, H, C.byref(self.cost.robot.desc), fields, nf, C.byref(gp),
                self.temperature, self.lr, opt_iters, _lib.stream_ptr()))
            self._noise.offset += opt_iters
            if opt_iters > 0:
                self.costs = self._cost_buf.view(P, S)
                self._weights = self._w_buf.view(P, S, 1, 1)
            return
        for it in range(opt_iters):
            self.costs = self._sample_and_eval(eps=None if eps is None else eps[it], **observation)
            self._update_distribution(self.costs, self.state_particles)

    def _can_run_fused(self, eps, observation):
        """mpb_stomp_run applies when nothing needs the host between the three kernels of an iteration."""
        c = self.cost
        return (eps is None and not observation and self.split.world == 1 and self._s_local <= self.SPLIT_THRESHOLD
                and hasattr(c, '_build') and hasattr(c, 'robot') and not c.has_extra_terms
                and self.d_state_opt == 2 * c.robot.desc.q_dim and type(self)._sample_and_eval is STOMP._sample_and_eval
                and type(self)._update_distribution is STOMP._update_distribution and self.state_particles.is_contiguous())

    def _sample_and_eval(self, eps=None, **observation):
        P, S = self.num_particles, self._s_local
        self.state_particles = self.sample(eps=eps)
        flat = self.state_particles.flatten(0, 1)
        if hasattr(self.cost, 'eval') and hasattr(self.cost, '_build'):
            costs = self.cost.eval(flat, out=self._cost_buf, **observation)
        else:
            costs = self._get_costs(flat, **observation)
        return costs.view(P, S)

    def _update_distribution(self, costs, traj_particles):
        P, S, H, D = self.num_particles, self._s_local, self.n_support_points, self.d_state_opt
        _lib.require_f32(costs, traj_particles)
        if self.split.world > 1 or S > self.SPLIT_THRESHOLD:
            # one problem with very many samples, or samples sharded over GPUs: partial records + fixed-order combine
            traj_particles = traj_particles.contiguous()
            r = split_softmax_update(costs, traj_particles, self._particle_means, self.temperature, self.lr, H, D,
                                     SigmaR=self.Sigma, split=self.split, S_global=self.num_samples, weights_out=self._w_buf)
            self._weights = r['weights'].view(P, S, 1, 1)
            return
        costs, traj_particles = costs.contiguous(), traj_particles.contiguous()     # named: temporaries must outlive the launch
        _lib.check(_lib.lib().mpb_softmax_update(_lib.ptr(costs), _lib.ptr(traj_particles),
                                                 _lib.ptr(self._particle_means), _lib.ptr(self._w_buf), None,
                                                 self.temperature, self.lr, _lib.ptr(self.Sigma), P, S, H, D,
                                                 _lib.stream_ptr()))
        self._weights = self._w_buf.view(P, S, 1, 1)

    def _calc_sample_weights(self, costs):
        return torch.softmax(-costs / self.temperature, dim=1)
