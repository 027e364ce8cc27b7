"""MPPI with the reference's constructor, attributes and return values (mp_baselines/planners/mppi.py).

    sample_and_eval     mppi.py:88-134   ->  mpb_mppi_rollout (control sampling + rollout + quadratic cost + IS dots)
                                             + mpb_cost_eval on the (state | control) rows (observation['cost'])
                                             + mpb_sum_f64 / mpb_mppi_finalize (batch-summed obstacle cost, quirk B2)
    update_controller   mppi.py:72-86    ->  mpb_softmax_partial + mpb_softmax_combine + mpb_softmax_weights
    _save_best          mppi.py:164-169  ->  (cmin, first argmin) from the same records; no host synchronisation

``sample_split``: optional motion_planning_baselines_b200.update.SampleSplit -- the N control samples of ONE problem
are sharded over the ranks of a process group; each rank rolls out its block and a single all-gather of the packed
records (plus one of the batch-sum scalar) makes every rank apply the identical update."""
import ctypes as C

import torch

from .. import _lib
from ..priors import get_multivar_gaussian_prior
from ..update import SampleSplit, split_softmax_update
from .base import MPPlanner

_CPU32 = dict(device='cpu', dtype=torch.float32)


class MPPI(MPPlanner):

    def __init__(self, system, num_ctrl_samples, rollout_steps, opt_iters, control_std=None, initial_mean=None,
                 step_size=1., temp=1., cov_prior_type='indep_ctrl', tensor_args=None, sample_split=None, seed=None,
                 **kwargs):
        super().__init__(name='MPPI', tensor_args=tensor_args)
        self.system = system
        self.state_dim = system.state_dim
        self.control_dim = system.control_dim
        self.rollout_steps = rollout_steps
        self.num_ctrl_samples = num_ctrl_samples
        self.opt_iters = opt_iters
        self.step_size = step_size
        self.temp = temp
        self._mean = torch.zeros(rollout_steps, self.control_dim, **self.tensor_args)
        self.control_std = control_std
        self.cov_prior_type = cov_prior_type
        self.weights = None
        self.ctrl_dist = get_multivar_gaussian_prior(control_std, rollout_steps, self.control_dim, Cov_type=cov_prior_type,
                                                     mu_init=self._mean, tensor_args=self.tensor_args)
        Cov_cpu = self.ctrl_dist.Cov.to(**_CPU32)
        self.Cov_inv = torch.stack([Cov_cpu[..., i].inverse() for i in range(self.control_dim)]).to(**self.tensor_args).contiguous()
        # one control_std for every dimension (the reference's 'const_ctrl' prior, gaussian.py:271-333): the C factors are
        # identical and the rollout kernel stages only one of them (checked once, bit for bit, on the host)
        L_cpu = self.ctrl_dist.scale_tril.detach().cpu()
        self._rollout_flags = 1 if all(torch.equal(L_cpu[0], L_cpu[i]) for i in range(1, self.control_dim)) else 0
        self.best_cost = torch.full((), float('inf'), **self.tensor_args)
        self.best_traj = torch.zeros(rollout_steps, self.state_dim, **self.tensor_args)
        self.split = sample_split or SampleSplit(world=1, rank=0)
        self._offset, self._n_local = self.split.local_slice(num_ctrl_samples)
        # in-kernel noise: one global Philox stream [C, N_glob, T]; a rank of a sample split draws its own samples
        self._noise = _lib.NoiseStream(seed, s_offset=self._offset, P_global=num_ctrl_samples)
        N, T, W = self._n_local, rollout_steps, self.state_dim + self.control_dim
        dev = self.tensor_args['device']
        self._xu = torch.empty(N, T, W, **self.tensor_args)
        self._quad = torch.empty(N, **self.tensor_args)
        self._isv = torch.empty(N, self.control_dim, **self.tensor_args)
        self._ext = torch.empty(N, **self.tensor_args)
        self._energy = torch.zeros(1, device=dev, dtype=torch.float64)
        self._scratch = torch.empty(1024, device=dev, dtype=torch.float64)
        self.costs = torch.empty(N, 1, **self.tensor_args)
        self._w_buf = torch.empty(1, N, **self.tensor_args)
        self.reset(initial_mean=initial_mean)

    def reset(self, initial_mean=None):
        if initial_mean is not None:
            self._mean = initial_mean.to(**self.tensor_args).contiguous().clone()
        else:
            self._mean = torch.zeros(self.rollout_steps, self.control_dim, **self.tensor_args)
        self.update_ctrl_dist()

    def update_ctrl_dist(self):
        self.ctrl_dist.update_means(self._mean)

    # ------------------------------------------------------------------ hot path
    def sample_and_eval(self, eps=None, **observation):
        """``eps``: optional injected noise [C, N_global, T] (one block per control dimension, as the reference's
        per-dimension MultivariateNormal.sample((N,)) draws them)."""
        N, T, C_, sd = self._n_local, self.rollout_steps, self.control_dim, self.state_dim
        lib, st = _lib.lib(), _lib.stream_ptr()
        eps_l, nd = None, None
        if eps is None:     # drawn inside the kernel (Philox keyed on the global element index)
            nd = C.byref(self._noise.next())
        else:
            _lib.require_f32(eps)
            assert eps.shape == (C_, self.num_ctrl_samples, T)
            eps_l = eps[:, self._offset:self._offset + N].contiguous()
        state0 = observation['state'].to(**self.tensor_args).contiguous()
        goal = observation.get('goal_state', self.system.goal_state).to(**self.tensor_args).contiguous()
        cw = self.system._c_weights
        # controls are sampled around ctrl_dist.mu (refreshed by update_ctrl_dist only), the IS term uses self._mean:
        # they differ after pop()/shift(), exactly as in the reference (mppi.py:68-70,125-128,171-178)
        _lib.check(lib.mpb_mppi_rollout_opt(
            _lib.ptr(self.ctrl_dist.scale_tril), _lib.ptr(self.Cov_inv), _lib.ptr(self._mean.contiguous()), _lib.ptr(self.ctrl_dist.mu.contiguous()),
            _lib.ptr(eps_l), nd, _lib.ptr(state0),
            _lib.ptr(goal), _lib.ptr(self.system.ctrl_min), _lib.ptr(self.system.ctrl_max), _lib.ptr(self._xu), _lib.ptr(self._quad),
            _lib.ptr(self._isv), N, T, C_, sd, float(self.system.dt), float(self.system.discount), float(cw['pos']),
            float(cw['ctrl']), float(cw['pos_T']), self._rollout_flags, st))
        cost = observation.get('cost', None)
        energy = None
        if cost is not None:
            cost.eval(self._xu, out=self._ext)                                   # obstacle cost of every rollout [N]
            _lib.check(lib.mpb_sum_f64(_lib.ptr(self._ext), N, _lib.ptr(self._energy), _lib.ptr(self._scratch), st))
            if self.split.world > 1:                                              # the batch sum runs over ALL samples
                self._energy = self.split.all_gather_cat(self._energy).sum(0, keepdim=True)
            energy = self._energy
        _lib.check(lib.mpb_mppi_finalize(_lib.ptr(self._quad), _lib.ptr(self._isv), _lib.ptr(energy), float(self.temp),
                                         _lib.ptr(self.costs), N, C_, st))
        self.state_trajectories = self._xu[..., :sd]
        return self._xu[..., sd:], self.state_trajectories, self.costs

    def update_controller(self, costs, U_sampled=None):
        """softmax over ALL samples + weighted-mean update of the control means.  ``U_sampled`` is accepted for
        signature parity; the kernel reads the control columns of the (state | control) rows in place."""
        N, T, C, sd = self._n_local, self.rollout_steps, self.control_dim, self.state_dim
        r = split_softmax_update(costs.reshape(1, N), self._xu.view(1, N, T, sd + C), self._mean.view(1, T, C), self.temp,
                                 self.step_size, T, sd + C, c0=sd, Dw=C, split=self.split, S_global=self.num_ctrl_samples,
                                 weights_out=self._w_buf)
        self.weights = r['weights'].view(N, 1)
        self._best = (r['best_cost'], r['best_idx'])
        self.update_ctrl_dist()

    def _save_best(self):
        """best_cost / best_traj across optimize() calls (first argmin, mppi.py:164-169), decided on the device."""
        cmin, idx = self._best
        better = cmin[0] < self.best_cost
        local = (idx[0].long() - self._offset).clamp(0, max(self._n_local - 1, 0))
        cand = self.state_trajectories[local] if self._n_local > 0 else torch.zeros_like(self.best_traj)
        if self.split.world > 1:
            owner = self.split.owner_of(idx[:1].long(), self.num_ctrl_samples)
            cand = self.split.all_gather_cat(cand.unsqueeze(0).contiguous()).index_select(0, owner)[0]
        self.best_cost = torch.where(better, cmin[0], self.best_cost)
        self.best_traj = torch.where(better, cand, self.best_traj)

    def optimize(self, opt_iters=None, eps=None, **observation):
        if opt_iters is None:
            opt_iters = self.opt_iters
        control_samples = state_trajectories = costs = None
        for it in range(opt_iters):
            control_samples, state_trajectories, costs = self.sample_and_eval(eps=None if eps is None else eps[it], **observation)
            self.update_controller(costs, control_samples)
            self._save_best()
        self._recent_control_samples = control_samples
        self._recent_state_trajectories = state_trajectories
        self._recent_weights = self.weights
        return control_samples, state_trajectories, costs

    def pop(self):
        action = self._mean[0, :].clone().detach()
        self.shift()
        return action

    def shift(self):
        self._mean = self._mean.roll(shifts=-1, dims=-1)      # sic: the reference rolls the control axis (quirk B11)
        self._mean[-1:] = 0.

    def get_recent_samples(self):
        return (self._recent_control_samples.detach().clone(), self._recent_state_trajectories.detach().clone(),
                self._recent_weights.detach().clone())

    def get_mean_controls(self):
        return self._mean

    def get_state_trajectories_rollout(self, controls=None, num_ctrl_samples=None, **observation):
        """Roll out given controls [n,T,C] (or the mean) -- zero noise through the same kernel."""
        T, C, sd = self.rollout_steps, self.control_dim, self.state_dim
        U = self._mean.unsqueeze(0) if controls is None else controls
        n = U.shape[0]
        lib, st, ta = _lib.lib(), _lib.stream_ptr(), self.tensor_args
        xu = torch.empty(n, T, sd + C, **ta)
        out = torch.empty(n, T, sd, **ta)
        zeros_eps = torch.zeros(C, 1, T, **ta)
        state0 = observation['state'].to(**ta).contiguous()
        goal = self.system.goal_state.to(**ta).contiguous()
        quad, isv = torch.empty(1, **ta), torch.empty(1, C, **ta)
        for i in range(n):          # utility path (the hot loop is sample_and_eval): mean = the given controls, eps = 0
            Ui = U[i].contiguous()
            _lib.check(lib.mpb_mppi_rollout(_lib.ptr(self.ctrl_dist.scale_tril), _lib.ptr(self.Cov_inv), _lib.ptr(Ui),
                                            _lib.ptr(zeros_eps), _lib.ptr(state0), _lib.ptr(goal), _lib.ptr(self.system.ctrl_min),
                                            _lib.ptr(self.system.ctrl_max), _lib.ptr(xu[i:i + 1]), _lib.ptr(quad), _lib.ptr(isv),
                                            1, T, C, sd, float(self.system.dt), float(self.system.discount), 0., 0., 0., st))
            out[i] = xu[i, :, :sd]
        return out
