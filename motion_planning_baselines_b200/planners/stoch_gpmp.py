"""Stoch-GPMP with the reference's constructor, attributes and return values
(mp_baselines/planners/stoch_gpmp.py), its iteration body replaced by the fused kernels:

    sample_and_eval        stoch_gpmp.py:244-265  ->  mpb_sample_gp + mpb_prior_matvec + mpb_cost_eval
    _update_distribution   stoch_gpmp.py:267-279  ->  mpb_softmax_update
    optimize               stoch_gpmp.py:281-309  ->  mpb_stoch_gpmp_iter per iteration (no host sync)
"""
import ctypes as C

import torch

from .. import _lib
from ..costs import build_gpmp2_cost_composite
from ..factors import GPFactor, MultiMPPrior, UnaryFactor
from ..update import split_softmax_update
from .base import OptimizationPlanner


class StochGPMP(OptimizationPlanner):
    SPLIT_THRESHOLD = 4096      # samples per particle above which the update is split over several CTAs (as STOMP)

    def __init__(self, robot=None, n_dof=None, n_support_points=None, num_particles_per_goal=None, opt_iters=None,
                 dt=None, start_state=None, step_size=1., multi_goal_states=None, initial_particle_means=None,
                 sigma_start_init=None, sigma_start_sample=None, sigma_goal_init=None, sigma_goal_sample=None,
                 sigma_gp_init=None, sigma_gp_sample=None, num_samples=2, temperature=1., seed=None,
                 noise_particle_offset=0, noise_particles_global=None, **kwargs):
        super().__init__(name='StochGPMP', n_dof=n_dof, n_support_points=n_support_points,
                         num_particles_per_goal=num_particles_per_goal, opt_iters=opt_iters, dt=dt,
                         start_state=start_state, initial_particle_means=initial_particle_means,
                         multi_goal_states=multi_goal_states, sigma_start_init=sigma_start_init,
                         sigma_goal_init=sigma_goal_init, sigma_gp_init=sigma_gp_init, pos_only=False, **kwargs)
        self.robot = robot
        # In-kernel noise (the reference draws inside MultiMPPrior.sample): one global Philox stream for the whole job.
        # A process that holds particles [noise_particle_offset, +num_particles) of noise_particles_global draws exactly
        # its slice, so results do not depend on how particles are sharded over GPUs.  seed: torch.initial_seed().
        P_glob = noise_particles_global if noise_particles_global is not None else noise_particle_offset + self.num_particles
        self._noise = _lib.NoiseStream(seed, p_offset=noise_particle_offset, P_global=P_glob)
        self._noise_init = _lib.NoiseStream(self._noise.seed, s_offset=noise_particle_offset, P_global=self.num_goals,
                                            offset=1 << 62)
        self.goal_directed = multi_goal_states is not None
        if self.goal_directed and self.num_goals != 1:
            raise NotImplementedError('Stoch-GPMP is single-goal (the reference cost breaks for > 1 goal, quirk B3)')
        self.num_samples = num_samples
        self.step_size = step_size
        self.temperature = temperature
        self.sigma_start_sample = sigma_start_sample
        self.sigma_goal_sample = sigma_goal_sample
        self.sigma_gp_sample = sigma_gp_sample
        self._mean = None
        self._weights = None
        self._sample_dist = None
        self._recent_state_particles = self._recent_control_particles = None
        self._has_recent = False
        self.costs = None
        self.free_flags = None

        self.cost = build_gpmp2_cost_composite(
            robot=robot, n_support_points=n_support_points, dt=dt, start_state=start_state.to(**self.tensor_args),
            multi_goal_states=None if multi_goal_states is None else multi_goal_states.to(**self.tensor_args),
            num_particles_per_goal=num_particles_per_goal, num_samples=num_samples, **kwargs)
        self.reset(initial_particle_means=initial_particle_means)

    # ------------------------------------------------------------------ setup (one-off, host side)
    def set_prior_factors(self):
        D, ta = self.d_state_opt, self.tensor_args
        self.start_prior_init = UnaryFactor(D, self.sigma_start_init, self.start_state, ta)
        self.gp_prior_init = GPFactor(self.n_dof, self.sigma_gp_init, self.dt, self.n_support_points - 1, ta)
        self.multi_goal_prior_init = [UnaryFactor(D, self.sigma_goal_init, g, ta) for g in self.multi_goal_states] \
            if self.goal_directed else []
        self.start_prior_sample = UnaryFactor(D, self.sigma_start_sample, self.start_state, ta)
        self.gp_prior_sample = GPFactor(self.n_dof, self.sigma_gp_sample, self.dt, self.n_support_points - 1, ta)
        self.multi_goal_prior_sample = [UnaryFactor(D, self.sigma_goal_sample, g, ta) for g in self.multi_goal_states] \
            if self.goal_directed else []

    def get_prior_dist(self, start_K, gp_K, goal_K, state_init, particle_means=None, goal_states=None, noise=None):
        return MultiMPPrior(self.n_support_points - 1, self.dt, 2 * self.n_dof, self.n_dof, start_K, gp_K, state_init,
                            K_g_inv=goal_K, means=particle_means, goal_states=goal_states, tensor_args=self.tensor_args,
                            noise=noise)

    def const_vel_trajectories(self, start_state, multi_goal_states):
        H, d = self.n_support_points, self.n_dof
        w = torch.arange(H, **self.tensor_args).view(1, H, 1)
        traj = torch.zeros(multi_goal_states.shape[0], self.num_particles_per_goal, H, self.d_state_opt, **self.tensor_args)
        pos = start_state[:d] * (H - w - 1) / (H - 1) + multi_goal_states[:, None, :d] * w / (H - 1)
        traj[..., :d] = pos.unsqueeze(1)
        traj[..., d:] = ((multi_goal_states[:, :d] - start_state[:d]) / (H * self.dt))[:, None, None, :]
        return traj

    def reset(self, start_state=None, multi_goal_states=None, initial_particle_means=None, eps_init=None):
        if start_state is not None:
            self.start_state = start_state.detach().clone().to(**self.tensor_args)
        if multi_goal_states is not None:
            self.multi_goal_states = multi_goal_states.detach().clone().to(**self.tensor_args)
        self.set_prior_factors()
        goal_K_init = self.multi_goal_prior_init[0].K if self.goal_directed else None
        goal_K_sample = self.multi_goal_prior_sample[0].K if self.goal_directed else None
        if initial_particle_means is not None:
            if isinstance(initial_particle_means, str) and initial_particle_means == 'const_vel':
                means = self.const_vel_trajectories(self.start_state, self.multi_goal_states)
            else:
                means = initial_particle_means.to(**self.tensor_args)
        else:
            init = self.get_prior_dist(self.start_prior_init.K, self.gp_prior_init.Q_inv[0], goal_K_init,
                                       self.start_state, goal_states=self.multi_goal_states, noise=self._noise_init)
            means = init.sample(self.num_particles_per_goal, eps=eps_init)
            del init
        self._particle_means = means.flatten(0, 1).contiguous().clone() if means.ndim == 4 else means.contiguous().clone()
        self._sample_dist = self.get_prior_dist(self.start_prior_sample.K, self.gp_prior_sample.Q_inv[0], goal_K_sample,
                                                self.start_state, particle_means=self._particle_means,
                                                goal_states=self.multi_goal_states, noise=self._noise)
        self.Sigma_inv = self._sample_dist.Sigma_inv
        # the reference's precision couples only equal dofs (<= 7 non-zeros per row): verified bit-exactly, once
        ok = C.c_int(0)
        _lib.check(_lib.lib().mpb_prior_dof_structured(_lib.ptr(self.Sigma_inv), self.n_support_points, self.n_dof,
                                                       C.byref(ok), _lib.stream_ptr()))
        self._sinv_structured = bool(ok.value)
        P, S, H, D = self.num_particles, self.num_samples, self.n_support_points, self.d_state_opt
        ta = self.tensor_args
        self._x_dm = None               # dof-major sample rows of the fused iteration (allocated on first use)
        self._x_dm_fresh = False        # True: _x_dm holds the latest samples and _state_samples has not been refreshed from it
        self.state_samples = torch.empty(P, S, H, D, **ta)
        self.costs = torch.empty(P, S, **ta)
        self._w_buf = torch.empty(P, S, **ta)
        self._is_vec = torch.empty(P, H * D, **ta)
        self.free_flags = torch.empty(P * S, device=ta['device'], dtype=torch.uint8)
        self.state_samples = self._sample_dist.sample(S, out=self.state_samples.view(P, S, H * D))

    # The fused iteration of the 7-dof arm keeps its sample rows DOF-MAJOR between its three kernels (include/mpb.h,
    # "dof-major sample rows"); the reference-layout [P, S, H, 2 dof] tensor is produced from them when somebody asks.
    @property
    def state_samples(self):
        if self._x_dm_fresh:
            P, S, H = self.num_particles, self.num_samples, self.n_support_points
            _lib.check(_lib.lib().mpb_traj_from_dof_major(_lib.ptr(self._x_dm), _lib.ptr(self._state_samples), P * S, H, self.n_dof,
                                                          _lib.stream_ptr()))
            self._x_dm_fresh = False
        return self._state_samples

    @state_samples.setter
    def state_samples(self, value):
        self._state_samples = value
        self._x_dm_fresh = False

    @property
    def _recent_control_samples(self):
        return self.state_samples[..., -self.n_dof:] if self._has_recent else None

    @_recent_control_samples.setter
    def _recent_control_samples(self, value):
        self._has_recent = value is not None

    @property
    def _recent_state_trajectories(self):
        return self.state_samples[..., :self.n_dof] if self._has_recent else None

    @_recent_state_trajectories.setter
    def _recent_state_trajectories(self, value):
        self._has_recent = value is not None

    def _use_dof_major(self, fields, nf):
        """The dof-major fused iteration: tcgen05 sampler with the mat-vec warp + a cost-kernel instance that reads the rows
        in place (7-dof chain, H = 64, primitive fields).  Bit-identical to the reference-layout iteration (tests/
        test_gpu_dof_major.py); MPB_X_DM=0 keeps the reference layout (A/B timing, tests)."""
        import os
        if os.environ.get('MPB_X_DM', '1') == '0':
            return False
        sd = self._sample_dist
        return (sd.scale_tril_kron_gen is not None and self._sinv_structured
                and bool(_lib.lib().mpb_cost_eval_dm_supported(C.byref(self.robot.desc), fields, nf, self.n_support_points)))

    def _dm_rows(self):
        if self._x_dm is None:
            self._x_dm = torch.empty(self.num_particles, self.num_samples, self.n_support_points * self.d_state_opt,
                                     **self.tensor_args)
        return self._x_dm

    # ------------------------------------------------------------------ hot path
    def _get_costs(self, **observation):
        """cost.eval + importance-sampling ratio term (stoch_gpmp.py:235-242) on self.state_samples."""
        P, S, M = self.num_particles, self.num_samples, self.n_support_points * self.d_state_opt
        self._prior_matvec(_lib.stream_ptr())
        self.cost.eval(self.state_samples, is_vec=self._is_vec, samples_per_particle=S, is_scale=self.temperature,
                       out=self.costs.view(-1), free_flag=self.free_flags, **observation)
        return self.costs

    def _prior_matvec(self, st):
        """is_vec[p] = Sigma_inv @ mu_p (first half of the IS term, stoch_gpmp.py:239-241)."""
        P, H, D = self.num_particles, self.n_support_points, self.d_state_opt
        if self._sinv_structured:
            _lib.check(_lib.lib().mpb_prior_matvec_dof(_lib.ptr(self.Sigma_inv), _lib.ptr(self._particle_means),
                                                       _lib.ptr(self._is_vec), P, H, self.n_dof, st))
        else:
            _lib.check(_lib.lib().mpb_prior_matvec(_lib.ptr(self.Sigma_inv), _lib.ptr(self._particle_means),
                                                   _lib.ptr(self._is_vec), P, H * D, 2 * D - 1, st))

    def sample_and_eval(self, eps=None, **observation):
        P, S, H, D = self.num_particles, self.num_samples, self.n_support_points, self.d_state_opt
        self._sample_dist.means = self._particle_means.view(P, -1)
        self.state_samples = self._sample_dist.sample(S, eps=eps, out=self.state_samples.view(P, S, H * D))
        costs = self._get_costs(**observation)
        d = self.n_dof
        return (self.state_samples[..., -d:], self.state_samples[..., :d],
                self._particle_means[..., -d:].clone(), self._particle_means[..., :d].clone(), costs)

    def _update_distribution(self, costs, traj_samples):
        P, S, H, D = self.num_particles, self.num_samples, self.n_support_points, self.d_state_opt
        costs, traj_samples = costs.contiguous(), traj_samples.contiguous()         # named: temporaries must outlive the launch
        if S > self.SPLIT_THRESHOLD:        # too many samples for one CTA's staging: partial records + fixed-order combine
            r = split_softmax_update(costs.view(P, S), traj_samples.view(P, S, H, D), self._particle_means, self.temperature,
                                     self.step_size, H, D, weights_out=self._w_buf, want_grad=True)
            self._weights = r['weights'].view(P, S, 1, 1)
            self._sample_dist.means = self._particle_means.view(P, -1)
            return r['grad']
        grad = torch.empty(P, H, D, **self.tensor_args)
        _lib.check(_lib.lib().mpb_softmax_update(_lib.ptr(costs), _lib.ptr(traj_samples),
                                                 _lib.ptr(self._particle_means), _lib.ptr(self._w_buf), _lib.ptr(grad),
                                                 self.temperature, self.step_size, None, P, S, H, D, _lib.stream_ptr()))
        self._weights = self._w_buf.view(P, S, 1, 1)
        self._sample_dist.means = self._particle_means.view(P, -1)
        return grad

    def optimize(self, opt_iters=None, debug=False, eps=None, **observation):
        """``eps``: optional injected noise, [opt_iters,S,P,M] or a list of [S,P,M] (parity runs)."""
        if opt_iters is None:
            opt_iters = self.opt_iters
        if observation.get('obstacle_spheres') is not None:
            raise NotImplementedError('per-call obstacle_spheres are not supported')
        P, S, H, D = self.num_particles, self.num_samples, self.n_support_points, self.d_state_opt
        M = H * D
        gp, fields, nf, _ = self.cost._build()
        if self.cost._extra is not None or S > self.SPLIT_THRESHOLD:
            # extra_costs (joint limits, ...) or a sample count beyond the single-CTA update: the staged path carries them
            return self._optimize_staged(opt_iters, eps, **observation)
        if opt_iters <= 0:
            return self._get_traj()
        lib = _lib.lib()
        pos_mean = vel_mean = None
        sd = self._sample_dist
        traj_out = None
        use_dm = eps is None and self._use_dof_major(fields, nf)
        for it in range(opt_iters):
            last = it == opt_iters - 1
            gen_path = eps is None and sd.scale_tril_kron_gen is not None
            if last:                    # the reference returns the pre-update particle means of the last iteration
                # default path: the kernels that touch the means anyway write both copies the reference API implies (the
                # pre-update means: K1's mat-vec warp; the returned clone: K3) -- no device copies of their own
                pre = torch.empty_like(self._particle_means) if gen_path else self._particle_means.clone()
                pos_mean, vel_mean = pre[..., :self.n_dof], pre[..., -self.n_dof:]       # the two halves are views of it
            if eps is None:
                nd = self._noise.next()
                if sd.scale_tril_kron_gen is not None:      # default: Blackwell sampler, noise drawn inside K1
                    if last:
                        traj_out = torch.empty_like(self._particle_means)
                    if use_dm:                              # sample rows stay dof-major between the three kernels
                        _lib.check(lib.mpb_stoch_gpmp_iter_kron_gen_dm(
                            _lib.ptr(sd.scale_tril_kron_gen), _lib.ptr(self.Sigma_inv), C.byref(nd),
                            _lib.ptr(self._particle_means), _lib.ptr(self._dm_rows()), _lib.ptr(self.costs), _lib.ptr(self._w_buf),
                            _lib.ptr(self._is_vec), _lib.ptr(self.free_flags), _lib.ptr(pre) if last else None,
                            _lib.ptr(traj_out) if last else None, P, S, H,
                            C.byref(self.robot.desc), fields, nf, C.byref(gp), self.temperature, self.step_size, _lib.stream_ptr()))
                        self._x_dm_fresh = True
                        continue
                    self._x_dm_fresh = False
                    _lib.check(lib.mpb_stoch_gpmp_iter_kron_gen_ex(
                        _lib.ptr(sd.scale_tril_kron_gen), _lib.ptr(self.Sigma_inv), int(self._sinv_structured), C.byref(nd),
                        _lib.ptr(self._particle_means), _lib.ptr(self.state_samples), _lib.ptr(self.costs), _lib.ptr(self._w_buf),
                        _lib.ptr(self._is_vec), _lib.ptr(self.free_flags), _lib.ptr(pre) if last else None,
                        _lib.ptr(traj_out) if last else None, P, S, H,
                        C.byref(self.robot.desc), fields, nf, C.byref(gp), self.temperature, self.step_size, _lib.stream_ptr()))
                    continue
                if sd.kron_tc_kind == 1:        # warp-MMA sampler, noise drawn inside K1
                    _lib.check(lib.mpb_stoch_gpmp_iter_kron_rng(
                        _lib.ptr(sd.scale_tril_kron_tc), _lib.ptr(self.Sigma_inv), int(self._sinv_structured), C.byref(nd),
                        _lib.ptr(self._particle_means), _lib.ptr(self.state_samples), _lib.ptr(self.costs), _lib.ptr(self._w_buf),
                        _lib.ptr(self._is_vec), _lib.ptr(self.free_flags), P, S, H,
                        C.byref(self.robot.desc), fields, nf, C.byref(gp), self.temperature, self.step_size, _lib.stream_ptr()))
                    continue
                e = _lib.philox_normal(nd, _lib.NOISE_SPM, (S, P, M), self.tensor_args['device'])
            else:
                e = eps[it]
            _lib.require_f32(e)
            assert e.shape == (S, P, M) and e.is_contiguous()
            self._x_dm_fresh = False
            if self._sample_dist.scale_tril_kron is not None:
                _lib.check(lib.mpb_stoch_gpmp_iter_kron(
                    _lib.ptr(self._sample_dist.scale_tril_kron), _lib.ptr(self._sample_dist.scale_tril_kron_tc),
                    self._sample_dist.kron_tc_kind, _lib.ptr(self.Sigma_inv), int(self._sinv_structured), _lib.ptr(e),
                    _lib.ptr(self._particle_means), _lib.ptr(self.state_samples), _lib.ptr(self.costs), _lib.ptr(self._w_buf),
                    _lib.ptr(self._is_vec), _lib.ptr(self.free_flags), P, S, H,
                    C.byref(self.robot.desc), fields, nf, C.byref(gp), self.temperature, self.step_size, _lib.stream_ptr()))
                continue
            _lib.check(lib.mpb_stoch_gpmp_iter(
                _lib.ptr(self._sample_dist.scale_tril), _lib.ptr(self._sample_dist.scale_tril_split),
                _lib.ptr(self.Sigma_inv), _lib.ptr(e),
                _lib.ptr(self._particle_means), _lib.ptr(self.state_samples), _lib.ptr(self.costs), _lib.ptr(self._w_buf),
                _lib.ptr(self._is_vec), _lib.ptr(self.free_flags), P, S, H,
                C.byref(self.robot.desc), fields, nf, C.byref(gp), self.temperature, self.step_size, _lib.stream_ptr()))
        self._weights = self._w_buf.view(P, S, 1, 1)
        self._sample_dist.means = self._particle_means.view(P, -1)
        self._has_recent = True         # _recent_control_samples / _recent_state_trajectories: views of state_samples, on demand
        self._recent_control_particles = vel_mean
        self._recent_state_particles = pos_mean
        self._recent_weights = self._weights
        return traj_out if traj_out is not None else self._get_traj()

    def _optimize_staged(self, opt_iters, eps=None, **observation):
        """optimize() through sample_and_eval + _update_distribution (composites with extra cost terms)."""
        P, S, M = self.num_particles, self.num_samples, self.n_support_points * self.d_state_opt
        if opt_iters <= 0:
            return self._get_traj()
        for it in range(opt_iters):
            e = eps[it] if eps is not None else None        # None: drawn on the device (Philox, self._noise)
            (self._recent_control_samples, self._recent_state_trajectories, self._recent_control_particles,
             self._recent_state_particles, costs) = self.sample_and_eval(eps=e, **observation)
            self._update_distribution(costs, self.state_samples)
        self._recent_weights = self._weights
        return self._get_traj()

    def step_fused(self):
        """One iteration on in-kernel noise as ONE C call (the three launches optimize() issues per iteration, without the
        bookkeeping around them: no returned clone, no pre-update copy).  Falls back to step_staged() where the fused entry
        points do not apply."""
        sd = self._sample_dist
        P, S, H = self.num_particles, self.num_samples, self.n_support_points
        gp, fields, nf, _ = self.cost._build()
        if sd.scale_tril_kron_gen is None or self.cost._extra is not None or S > self.SPLIT_THRESHOLD:
            return self.step_staged(None)
        lib, st = _lib.lib(), _lib.stream_ptr()
        nd = sd.noise.next()
        if self._use_dof_major(fields, nf):
            _lib.check(lib.mpb_stoch_gpmp_iter_kron_gen_dm(
                _lib.ptr(sd.scale_tril_kron_gen), _lib.ptr(self.Sigma_inv), C.byref(nd), _lib.ptr(self._particle_means),
                _lib.ptr(self._dm_rows()), _lib.ptr(self.costs), _lib.ptr(self._w_buf), _lib.ptr(self._is_vec),
                _lib.ptr(self.free_flags), None, None, P, S, H, C.byref(self.robot.desc), fields, nf, C.byref(gp),
                self.temperature, self.step_size, st))
            self._x_dm_fresh = True
            return
        self._x_dm_fresh = False
        _lib.check(lib.mpb_stoch_gpmp_iter_kron_gen_ex(
            _lib.ptr(sd.scale_tril_kron_gen), _lib.ptr(self.Sigma_inv), int(self._sinv_structured), C.byref(nd),
            _lib.ptr(self._particle_means), _lib.ptr(self._state_samples), _lib.ptr(self.costs), _lib.ptr(self._w_buf),
            _lib.ptr(self._is_vec), _lib.ptr(self.free_flags), None, None, P, S, H, C.byref(self.robot.desc), fields, nf,
            C.byref(gp), self.temperature, self.step_size, st))

    def step_staged(self, eps=None, events=None):
        """One iteration as four separate C-ABI calls (same kernels as mpb_stoch_gpmp_iter*); ``events`` is an
        optional list of 5 torch.cuda.Event recorded around the stages (bench.py per-kernel timing).  ``eps`` None:
        the noise is drawn inside K1 (default structured sampler) -- the way optimize() runs without injected noise."""
        P, S, H, D = self.num_particles, self.num_samples, self.n_support_points, self.d_state_opt
        M = H * D
        lib, st = _lib.lib(), _lib.stream_ptr()
        gp, fields, nf, _ = self.cost._build()
        rec = (lambda i: events[i].record()) if events is not None else (lambda i: None)
        rec(0)
        split = self._sample_dist.scale_tril_split
        mv_fused = False
        if eps is None and self._use_dof_major(fields, nf):
            # the three kernels of mpb_stoch_gpmp_iter_kron_gen_dm, one C call each
            nd = self._sample_dist.noise.next()
            xdm = self._dm_rows()
            _lib.check(lib.mpb_sample_gp_kron_gen_dm(_lib.ptr(self._sample_dist.scale_tril_kron_gen), _lib.ptr(self._particle_means),
                                                     C.byref(nd), _lib.ptr(xdm), P, S, H, self.n_dof,
                                                     _lib.ptr(self.Sigma_inv), _lib.ptr(self._is_vec), None, st))
            rec(1)
            rec(2)
            _lib.check(lib.mpb_cost_eval_dm(_lib.ptr(xdm), P * S, H, C.byref(self.robot.desc), fields, nf,
                                            C.byref(gp), _lib.ptr(self._is_vec), S, self.temperature, _lib.ptr(self.costs), None,
                                            _lib.ptr(self.free_flags), st))
            rec(3)
            _lib.check(lib.mpb_softmax_update_dm(_lib.ptr(self.costs), _lib.ptr(xdm), _lib.ptr(self._particle_means),
                                                 _lib.ptr(self._w_buf), None, self.temperature, self.step_size, None, P, S, H, D, st))
            rec(4)
            self._x_dm_fresh = True
            return
        self._x_dm_fresh = False
        if eps is None and self._sample_dist.scale_tril_kron_gen is not None and self._sinv_structured:
            # the tcgen05 sampler computes Sigma^-1 mu on one extra warp per CTA (as mpb_stoch_gpmp_iter_kron_gen runs it)
            nd = self._sample_dist.noise.next()
            _lib.check(lib.mpb_sample_gp_kron_gen_mv(_lib.ptr(self._sample_dist.scale_tril_kron_gen), _lib.ptr(self._particle_means),
                                                     C.byref(nd), _lib.ptr(self.state_samples), P, S, H, self.n_dof,
                                                     _lib.ptr(self.Sigma_inv), _lib.ptr(self._is_vec), None, st))
            mv_fused = True
        elif eps is None:
            self._sample_dist.means = self._particle_means.view(P, -1)
            self._sample_dist.sample(S, out=self.state_samples.view(P, S, M))
        elif self._sample_dist.kron_tc_kind == 2:
            _lib.check(lib.mpb_sample_gp_kron_umma(_lib.ptr(self._sample_dist.scale_tril_kron_tc), _lib.ptr(self._particle_means),
                                                   _lib.ptr(eps), _lib.ptr(self.state_samples), P, S, H, self.n_dof, st))
        elif self._sample_dist.kron_tc_kind == 1:
            _lib.check(lib.mpb_sample_gp_kron_tc(_lib.ptr(self._sample_dist.scale_tril_kron_tc), _lib.ptr(self._particle_means),
                                                 _lib.ptr(eps), _lib.ptr(self.state_samples), P, S, H, self.n_dof, st))
        elif self._sample_dist.scale_tril_kron is not None:
            _lib.check(lib.mpb_sample_gp_kron(_lib.ptr(self._sample_dist.scale_tril_kron), _lib.ptr(self._particle_means),
                                              _lib.ptr(eps), _lib.ptr(self.state_samples), P, S, H, self.n_dof, st))
        elif split is not None:
            _lib.check(lib.mpb_sample_gp_tc(_lib.ptr(split[0]), _lib.ptr(split[1]), _lib.ptr(self._particle_means), _lib.ptr(eps),
                                            _lib.ptr(self.state_samples), P, S, M, st))
        else:
            _lib.check(lib.mpb_sample_gp(_lib.ptr(self._sample_dist.scale_tril), _lib.ptr(self._particle_means), _lib.ptr(eps),
                                         _lib.ptr(self.state_samples), P, S, M, st))
        rec(1)
        if not mv_fused:
            self._prior_matvec(st)
        rec(2)
        _lib.check(lib.mpb_cost_eval(_lib.ptr(self.state_samples), P * S, H, C.byref(self.robot.desc), fields, nf,
                                     C.byref(gp), _lib.ptr(self._is_vec), S, self.temperature, _lib.ptr(self.costs), None,
                                     _lib.ptr(self.free_flags), st))
        rec(3)
        _lib.check(lib.mpb_softmax_update(_lib.ptr(self.costs), _lib.ptr(self.state_samples), _lib.ptr(self._particle_means),
                                          _lib.ptr(self._w_buf), None, self.temperature, self.step_size, None, P, S, H, D, st))
        rec(4)

    def get_recent_samples(self):
        if self._recent_state_particles is None:
            raise _lib.MpbError('get_recent_samples(): no optimize() iteration has run yet')
        return (self._recent_state_trajectories.detach().clone(), self._recent_state_particles.detach().clone(),
                self._recent_control_samples.detach().clone(), self._recent_control_particles.detach().clone(),
                self._recent_weights.detach().clone())

    def sample_trajectories(self, num_samples_per_particle):
        self._sample_dist.means = self._particle_means.view(self.num_particles, -1)
        samples = self._sample_dist.sample(num_samples_per_particle)
        return samples[..., :self.n_dof], samples[..., -self.n_dof:]
