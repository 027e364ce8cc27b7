"""Cost objects with the reference's names, constructor arguments and call surface
(mp_baselines/planners/costs/cost_functions.py), evaluated by ONE fused sm_100a kernel
(mpb_cost_eval, csrc/cost_eval.cu) instead of a chain of eager torch ops.

    Cost.eval(trajs [B,H,D] | [N,B,H,D], **obs) -> [B] (or [N*B])      cost_functions.py:30-35
    CostComposite(robot, n_support_points, cost_list, weights_cost_l)   cost_functions.py:56-105
    CostCollision(robot, n_support_points, field=, sigma_coll=)         cost_functions.py:147-189
    CostGP(robot, n_support_points, start_state, dt, sigma_params)      cost_functions.py:234-289
    CostGoalPrior(robot, n_support_points, multi_goal_states=, ...)     cost_functions.py:488-536
    CostGPTrajectory(robot, n_support_points, dt, sigma_gp=)            cost_functions.py:317-357
    CostGPTrajectoryPositionOnlyWrapper(...)                            cost_functions.py:360-368
    CostSmoothnessCHOMP(robot, n_support_points)                        cost_functions.py:371-390
    CostJointLimits(robot, n_support_points, eps=)                      cost_functions.py:393-429
    build_gpmp2_cost_composite(...)                                      gpmp2.py:23-89

There is no CPU implementation behind these classes: tensors must live on the CUDA device.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .fields import Field


class Cost:
    def __init__(self, robot, n_support_points, tensor_args=None, **kwargs):
        self.robot = robot
        self.n_dof = robot.q_dim
        self.dim = 2 * self.n_dof
        self.n_support_points = n_support_points
        self.tensor_args = tensor_args if tensor_args is not None else robot.tensor_args

    def set_cost_factors(self):
        pass

    def __call__(self, trajs, **kwargs):
        return self.eval(trajs, **kwargs)

    def eval(self, trajs, **kwargs):
        return CostComposite(self.robot, self.n_support_points, [self], tensor_args=self.tensor_args).eval(trajs, **kwargs)

    def get_linear_system(self, trajs, **kwargs):
        raise NotImplementedError


class CostCollision(Cost):
    def __init__(self, robot, n_support_points, field=None, sigma_coll=None, **kwargs):
        super().__init__(robot, n_support_points, **kwargs)
        if field is not None and not isinstance(field, Field):
            raise _lib.MpbError('CostCollision needs a motion_planning_baselines_b200.fields.Field '
                                '(CollisionField, SelfCollisionField or WorkspaceBoundaryField)')
        self.field = field
        if field is not None and field.robot is None:
            field.bind_robot(robot)                         # compute_cost() needs the sphere table (plug point 2)
        self.sigma_coll = sigma_coll
        self.inv_sigma2 = 1. / (sigma_coll ** 2)          # FieldFactor.K (field_factor.py:15)


class CostGP(Cost):
    def __init__(self, robot, n_support_points, start_state, dt, sigma_params, **kwargs):
        super().__init__(robot, n_support_points, **kwargs)
        self.start_state = start_state.to(**self.tensor_args).contiguous()
        assert self.start_state.numel() == self.dim, 'start_state must be [2*dof] (position | zero velocity)'
        self.dt = dt
        self.sigma_start = sigma_params['sigma_start']
        self.sigma_gp = sigma_params['sigma_gp']
        # fp32 constants exactly as UnaryFactor.K / GPFactor.calc_Q_inv build them
        one = torch.ones((), dtype=torch.float32)
        self.k_start = float(one / self.sigma_start ** 2)
        qc = one / self.sigma_gp ** 2
        self.q11 = float(12. * (dt ** -3.) * qc)
        self.q12 = float(-6. * (dt ** -2.) * qc)
        self.q22 = float(4. * (dt ** -1.) * qc)


class CostGoalPrior(Cost):
    def __init__(self, robot, n_support_points, multi_goal_states=None, num_particles_per_goal=None,
                 num_samples=None, sigma_goal_prior=None, **kwargs):
        super().__init__(robot, n_support_points, **kwargs)
        self.multi_goal_states = multi_goal_states
        self.num_goals = multi_goal_states.shape[0]
        if self.num_goals != 1:
            # the reference reshapes to [num_goals, ...] and breaks for > 1 goal
            raise NotImplementedError('CostGoalPrior supports a single goal (as the reference effectively does)')
        self.goal_state = multi_goal_states.reshape(-1)[:self.dim].to(**self.tensor_args).contiguous()
        self.num_particles_per_goal = num_particles_per_goal
        self.num_samples = num_samples
        self.sigma_goal_prior = sigma_goal_prior
        self.k_goal = float(torch.ones((), dtype=torch.float32) / sigma_goal_prior ** 2)


class CostGPTrajectory(Cost):
    """GP-prior smoothness of a whole trajectory without the start term (cost_functions.py:317-357)."""

    def __init__(self, robot, n_support_points, dt, sigma_gp=None, **kwargs):
        super().__init__(robot, n_support_points, **kwargs)
        self.dt = dt
        self.sigma_gp = sigma_gp
        qc = torch.ones((), dtype=torch.float32) / sigma_gp ** 2
        self.q11 = float(12. * (dt ** -3.) * qc)
        self.q12 = float(-6. * (dt ** -2.) * qc)
        self.q22 = float(4. * (dt ** -1.) * qc)


class CostGPTrajectoryPositionOnlyWrapper(CostGPTrajectory):
    """Position-only trajectories [B,H,d]: velocities by central differences (zero at both ends), then the
    GP-trajectory cost (cost_functions.py:360-368; ``finite_difference_vector`` is external -- semantics as in
    oracle/ref_shim)."""

    def eval(self, trajs, **observation):
        vel = torch.zeros_like(trajs)
        vel[..., 1:-1, :] = (trajs[..., 2:, :] - trajs[..., :-2, :]) / (2 * self.dt)
        full = torch.cat((trajs, vel), dim=-1)
        inner = CostGPTrajectory(self.robot, self.n_support_points, self.dt, sigma_gp=self.sigma_gp, tensor_args=self.tensor_args)
        return inner.eval(full, **observation)


class CostSmoothnessCHOMP(Cost):
    """x[:, :, j]^T R x[:, :, j] with CHOMP's finite-difference precision R -> [B, D] (cost_functions.py:371-390).
    Like the reference's it returns one value per state column, so it cannot be a member of a composite."""

    def __init__(self, robot, n_support_points, **kwargs):
        super().__init__(robot, n_support_points, **kwargs)
        self.dt = robot.dt
        from .planners.chomp import CHOMP
        self.Sigma_inv = CHOMP._get_R_mat(dt=self.dt, n_support_points=n_support_points, tensor_args=self.tensor_args)

    def eval(self, trajs, **observation):
        _lib.require_f32(trajs)
        x = trajs.contiguous()
        B, H, D = x.shape
        out = torch.empty(B, D, device=x.device, dtype=torch.float32)
        _lib.check(_lib.lib().mpb_smoothness_cost(_lib.ptr(x), _lib.ptr(self.Sigma_inv), _lib.ptr(out), B, H, D, _lib.stream_ptr()))
        return out


class CostJointLimits(Cost):
    """Squared violation of the joint limits shrunk by eps (cost_functions.py:393-429).  As in the reference the
    value is SUMMED OVER THE WHOLE BATCH into one 0-dim tensor (its ``.sum(-1)`` runs over the flat list of violating
    entries), which a composite then adds to every trajectory."""

    def __init__(self, robot, n_support_points, eps=np.deg2rad(3), **kwargs):
        super().__init__(robot, n_support_points, **kwargs)
        self.eps = float(eps)

    def eval(self, trajs, **observation):
        assert trajs.ndim == 3
        comp = CostComposite(self.robot, self.n_support_points, [self], tensor_args=self.tensor_args)
        return comp._eval_full(trajs)[1]


class CostComposite(Cost):
    def __init__(self, robot, n_support_points, cost_list, weights_cost_l=None, **kwargs):
        super().__init__(robot, n_support_points, **kwargs)
        self.cost_l = list(cost_list)
        self.weight_cost_l = weights_cost_l if weights_cost_l is not None else [1.0] * len(self.cost_l)
        self._build()

    def _build(self, unit_weights=False):
        gp = _lib.GPDesc(enabled=0, has_goal=0, w_gp=1.0, w_goal=1.0)
        fields = []
        self._extra = None
        self._w_jl = 0.0
        self._keepalive = []
        order = []                                          # kernel term index of every cost_l entry
        n_gp = n_goal = 0
        for cost, w in zip(self.cost_l, self.weight_cost_l):
            w = 1.0 if unit_weights else float(w)
            if isinstance(cost, CostGP):
                assert n_gp == 0, 'one CostGP per composite'
                n_gp += 1
                gp.enabled, gp.dt, gp.k_start = 1, cost.dt, cost.k_start
                gp.q11, gp.q12, gp.q22, gp.w_gp = cost.q11, cost.q12, cost.q22, w
                gp.start_state = cost.start_state.data_ptr()
                self._keepalive.append(cost.start_state)
                order.append(('gp', 0))
            elif isinstance(cost, CostGoalPrior):
                assert n_goal == 0, 'one CostGoalPrior per composite'
                n_goal += 1
                gp.has_goal, gp.k_goal, gp.w_goal = 1, cost.k_goal, w
                gp.goal_state = cost.goal_state.data_ptr()
                self._keepalive.append(cost.goal_state)
                order.append(('goal', 0))
            elif isinstance(cost, CostCollision):
                if cost.field is None:
                    order.append(('none', 0))
                    continue
                order.append(('field', len(fields)))
                fields.append(cost.field.desc(weight=w, inv_sigma2=cost.inv_sigma2))
            elif isinstance(cost, CostGPTrajectory):
                ex = self._extra or _lib.ExtraCostDesc()
                assert not ex.gp_traj_enabled, 'one CostGPTrajectory per composite'
                if gp.enabled and abs(gp.dt - cost.dt) > 1e-6 * abs(cost.dt):
                    raise NotImplementedError('CostGP and CostGPTrajectory must share dt')
                gp.dt = cost.dt
                ex.gp_traj_enabled, ex.t11, ex.t12, ex.t22, ex.w_gp_traj = 1, cost.q11, cost.q12, cost.q22, w
                self._extra = ex
                order.append(('gp_traj', 0))
            elif isinstance(cost, CostJointLimits):
                ex = self._extra or _lib.ExtraCostDesc()
                assert not ex.jl_enabled, 'one CostJointLimits per composite'
                ex.jl_enabled, ex.jl_eps, ex.w_jl = 1, cost.eps, w
                ex.q_min, ex.q_max = self.robot.q_min.data_ptr(), self.robot.q_max.data_ptr()
                self._extra = ex
                self._w_jl = w
                order.append(('jl', 0))
            else:
                raise NotImplementedError(f'{type(cost).__name__} has no fused implementation (CostSmoothnessCHOMP returns '
                                          f'[B,D] and cannot be summed into a composite in the reference either)')
        if gp.has_goal and not gp.enabled:
            raise NotImplementedError('CostGoalPrior without CostGP is not supported by the fused kernel')
        if len(fields) > _lib.MPB_MAX_FIELDS:
            raise NotImplementedError(f'at most {_lib.MPB_MAX_FIELDS} collision fields per composite')
        arr = (_lib.FieldDesc * max(1, len(fields)))(*fields)
        n_head = int(gp.enabled) + int(gp.has_goal)
        term_index = []
        for kind, i in order:
            term_index.append({'gp': 0, 'goal': 1, 'field': n_head + i, 'none': -1, 'gp_traj': n_head + len(fields),
                               'jl': -2}[kind])
        return gp, arr, len(fields), term_index

    @property
    def has_extra_terms(self):
        self._build()
        return self._extra is not None

    def _flatten(self, trajs):
        assert trajs.ndim in (3, 4)
        if trajs.ndim == 4:
            trajs = trajs.reshape(-1, *trajs.shape[2:])
        _lib.require_f32(trajs)
        if trajs.shape[-1] != self.dim or trajs.shape[-2] != self.n_support_points:
            raise _lib.MpbError(f'trajectories must be [..., {self.n_support_points}, {self.dim}], got {tuple(trajs.shape)}')
        return trajs.contiguous()

    def eval(self, trajs, trajs_interpolated=None, return_invidual_costs_and_weights=False,
             is_vec=None, samples_per_particle=1, is_scale=0.0, out=None, free_flag=None, **kwargs):
        """Reference signature plus optional fused extras (IS term, collision-free flags).

        ``trajs_interpolated`` is accepted and has NO effect, exactly as in the reference: its composite forwards the
        q_pos / H_positions of ``trajs`` to every term (cost_functions.py:71,85), FieldFactor.get_error uses those
        (field_factor.py:31-39), so the collision term is evaluated on the support points either way (verified by
        running the reference: oracle/make_golden_next.py, tests/golden/extra_costs_*.npz)."""
        if trajs_interpolated is not None and trajs_interpolated.shape[0] != trajs.reshape(-1, *trajs.shape[-2:]).shape[0]:
            raise _lib.MpbError('trajs_interpolated must have the batch size of trajs')
        if kwargs.get('obstacle_spheres') is not None:
            raise NotImplementedError('per-call obstacle_spheres are not supported')
        if trajs.requires_grad and torch.is_grad_enabled() and not return_invidual_costs_and_weights:
            # the reference CHOMP loop differentiates the cost object: costs.sum().backward() (chomp.py:134-139)
            if is_vec is not None:
                raise NotImplementedError('the importance-sampling term has no fused gradient')
            return _CostEvalGrad.apply(trajs, self)
        res = self._eval_full(trajs, return_terms=return_invidual_costs_and_weights, is_vec=is_vec,
                              samples_per_particle=samples_per_particle, is_scale=is_scale, out=out, free_flag=free_flag)
        if return_invidual_costs_and_weights:
            return res[2], self.weight_cost_l
        return res[0]

    def _eval_full(self, trajs, return_terms=False, is_vec=None, samples_per_particle=1, is_scale=0.0, out=None,
                   free_flag=None):
        """-> (total [B], joint-limit batch sum (0-dim) | None, per-cost list | None)."""
        x = self._flatten(trajs)
        B = x.shape[0]
        ta = dict(device=x.device, dtype=torch.float32)
        if B == 0 and not return_terms:
            return (out if out is not None else torch.empty(0, **ta)), torch.zeros((), **ta), None
        gp, fields, nf, term_index = self._build(unit_weights=return_terms)
        extra = self._extra
        cost = out if out is not None else torch.empty(B, **ta)
        n_terms = int(gp.enabled) + int(gp.has_goal) + nf + int(bool(extra is not None and extra.gp_traj_enabled))
        terms = torch.empty(max(n_terms, 1), B, **ta) if return_terms else None
        jl = torch.empty(B, **ta) if extra is not None and extra.jl_enabled else None
        _lib.check(_lib.lib().mpb_cost_eval_ex(
            _lib.ptr(x), B, self.n_support_points, C.byref(self.robot.desc), fields, nf, C.byref(gp),
            _lib.ptr(is_vec), samples_per_particle, is_scale,
            _lib.ptr(cost), _lib.ptr(terms), _lib.ptr(free_flag),
            C.byref(extra) if extra is not None else None, _lib.ptr(jl), _lib.stream_ptr()))
        jl_sum = None
        if jl is not None:
            # the reference's joint-limit term is one scalar for the whole batch, added to every trajectory
            acc = torch.empty(1, device=x.device, dtype=torch.float64)
            scratch = torch.empty(1024, device=x.device, dtype=torch.float64)
            _lib.check(_lib.lib().mpb_sum_f64(_lib.ptr(jl), B, _lib.ptr(acc), _lib.ptr(scratch), _lib.stream_ptr()))
            jl_sum = acc[0].to(torch.float32)
            if not return_terms:
                cost.add_(jl_sum * self._w_jl)
        per_cost = None
        if return_terms:
            zero = torch.zeros(B, **ta)
            per_cost = [jl_sum if i == -2 else (terms[i] if i >= 0 else zero) for i in term_index]
        return cost, jl_sum, per_cost

    def grad(self, trajs, grad_out=None):
        """d (sum_b grad_out[b] * cost[b]) / d trajs, analytic (``mpb_cost_grad``): what autograd through FK + SDF
        gives the reference (chomp.py:139).  trajs [B,H,D] (or [N,B,H,D]); grad_out [B] | None (ones)."""
        x = self._flatten(trajs.detach())
        B = x.shape[0]
        gp, fields, nf, _ = self._build()
        extra = self._extra
        if extra is not None and extra.gp_traj_enabled and not gp.enabled:
            gp.dt = [c for c in self.cost_l if isinstance(c, CostGPTrajectory)][0].dt
        g = torch.empty_like(x)
        go = None if grad_out is None else grad_out.reshape(-1).contiguous().float()
        gsum = None
        if extra is not None and extra.jl_enabled:     # the joint-limit scalar is added to EVERY trajectory of the batch
            gsum = go.sum().reshape(1) if go is not None else torch.full((1,), float(B), device=x.device, dtype=torch.float32)
        _lib.check(_lib.lib().mpb_cost_grad(
            _lib.ptr(x), B, self.n_support_points, C.byref(self.robot.desc), fields, nf, C.byref(gp),
            C.byref(extra) if extra is not None else None, _lib.ptr(go), _lib.ptr(gsum), _lib.ptr(g), _lib.stream_ptr()))
        return g.view(trajs.shape)

    @staticmethod
    def interpolation_weights(n_interpolated_points):
        """(n, host float array [n+1]) for mpb_gpmp2_linearize_ex: the weights of ``interpolate_points_v1`` (external;
        linear up-sampling with n extra points per segment, see oracle/costs.py)."""
        n = int(n_interpolated_points or 0)
        if n <= 0:
            return 0, None
        if n > _lib.MPB_MAX_INTERP:
            raise _lib.MpbError(f'at most {_lib.MPB_MAX_INTERP} interpolated points per segment')
        w = torch.linspace(0, 1, n + 2, dtype=torch.float32)[:-1]
        return n, (C.c_float * (n + 1))(*[float(v) for v in w])

    def linearize_collision(self, trajs, n_interpolated_points=None):
        """-> err [n_fields,B,H], hobs [n_fields,B,H,d]: collision errors and H_obst = -d err/d q of every
        waypoint (mpb_gpmp2_linearize; field_factor.py:41-57 without autograd).  With ``n_interpolated_points`` the
        Jacobian is that of the up-sampled trajectory's summed error (cost_functions.py:115-119)."""
        x = self._flatten(trajs)
        B, H, d = x.shape[0], self.n_support_points, self.n_dof
        gp, fields, nf, _ = self._build()
        err = torch.zeros(max(nf, 1), B, H, device=x.device, dtype=torch.float32)
        hobs = torch.zeros(max(nf, 1), B, H, d, device=x.device, dtype=torch.float32)
        n, w = self.interpolation_weights(n_interpolated_points)
        _lib.check(_lib.lib().mpb_gpmp2_linearize_ex(_lib.ptr(x), B, H, C.byref(self.robot.desc), fields, nf,
                                                     _lib.ptr(err), _lib.ptr(hobs), None, n, w, _lib.stream_ptr()))
        return err[:nf], hobs[:nf]

    def get_linear_system(self, trajs, n_interpolated_points=None, **kwargs):
        """Dense (A [B,rows,N], b [B,rows,1], K [B,rows,rows]) with the reference's row order
        (cost_functions.py:107-144,191-231,291-314,538-554).  Kept for API parity and diagnostics: the collision
        rows come from the analytic-Jacobian kernel, the (constant) GP rows are laid out with device tensor ops.
        GPMP2._step never calls this -- it solves the block-tridiagonal system without materialising A or K."""
        x = self._flatten(trajs)
        B, H, D, d = x.shape[0], self.n_support_points, self.dim, self.n_dof
        N, ta = H * D, dict(device=x.device, dtype=torch.float32)
        err, hobs = self.linearize_collision(x, n_interpolated_points=n_interpolated_points)
        eye = torch.eye(D, **ta)
        As, bs, Ks = [], [], []
        fi = 0
        for cost in self.cost_l:
            if isinstance(cost, CostGP):
                Phi = torch.eye(D, **ta)
                Phi[:d, d:] = torch.eye(d, **ta) * cost.dt
                Q = torch.zeros(D, D, **ta)
                Q[:d, :d], Q[:d, d:], Q[d:, :d], Q[d:, d:] = (torch.eye(d, **ta) * v for v in (cost.q11, cost.q12, cost.q12, cost.q22))
                A = torch.zeros(B, N, N, **ta)
                A[:, :D, :D] = eye
                r = torch.arange(D, N, device=x.device)
                A[:, r, r] = -1.0
                A[:, D:, :-D] += torch.kron(torch.eye(H - 1, **ta), Phi)
                b = torch.cat((cost.start_state - x[:, 0], (x[:, 1:] - x[:, :-1] @ Phi.t()).reshape(B, -1)), dim=1).unsqueeze(-1)
                K = torch.block_diag(eye * cost.k_start, *([Q] * (H - 1))).expand(B, N, N)
            elif isinstance(cost, CostGoalPrior):
                A = torch.zeros(B, D, N, **ta)
                A[:, :, -D:] = eye
                b = (cost.goal_state - x[:, -1]).unsqueeze(-1)
                K = (eye * cost.k_goal).expand(B, D, D)
            elif isinstance(cost, CostCollision):
                if cost.field is None:
                    continue
                A = torch.zeros(B, H - 1, H, D, **ta)
                t = torch.arange(H - 1, device=x.device)
                A[:, t, t + 1, :d] = hobs[fi][:, 1:]
                A = A.reshape(B, H - 1, N)
                b = err[fi][:, 1:].unsqueeze(-1)
                K = (torch.eye(H - 1, **ta) * cost.inv_sigma2).expand(B, H - 1, H - 1)
                fi += 1
            elif isinstance(cost, (CostGPTrajectory, CostJointLimits)):
                continue            # their get_linear_system returns nothing in the reference (cost_functions.py:356,428)
            else:
                raise NotImplementedError(type(cost).__name__)
            As.append(A), bs.append(b), Ks.append(K)
        A, b = torch.cat(As, dim=1), torch.cat(bs, dim=1)
        rows = A.shape[1]
        K = torch.zeros(B, rows, rows, **ta)
        o = 0
        for k in Ks:
            n = k.shape[1]
            K[:, o:o + n, o:o + n] = k
            o += n
        return A, b, K

    def collision_free(self, trajs):
        """bool [B]: every collision hinge term of the trajectory (waypoints 1..H-1) is exactly 0."""
        x = self._flatten(trajs)
        flags = torch.empty(x.shape[0], device=x.device, dtype=torch.uint8)
        self._eval_full(x, free_flag=flags)
        return flags.bool()


class _CostEvalGrad(torch.autograd.Function):
    """CostComposite.eval with an analytic backward (plug point 1 of SURVEY 8b: a cost object handed to the reference
    CHOMP must survive ``costs.sum().backward()``)."""

    @staticmethod
    def forward(ctx, trajs, composite):
        ctx.composite = composite
        ctx.save_for_backward(trajs)
        return composite._eval_full(trajs.detach())[0]

    @staticmethod
    def backward(ctx, grad_out):
        (trajs,) = ctx.saved_tensors
        return ctx.composite.grad(trajs, grad_out), None


def build_gpmp2_cost_composite(robot=None, n_support_points=None, dt=None, start_state=None, multi_goal_states=None,
                               num_particles_per_goal=None, collision_fields=None, extra_costs=[],
                               sigma_start=1e-5, sigma_gp=1e-2, sigma_coll=1e-5, sigma_goal_prior=1e-5,
                               num_samples: int = 64, tensor_args=None, **kwargs):
    """Same defaults and wiring as mp_baselines/planners/gpmp2.py:23-89."""
    cost_func_list = []
    start_state_zero_vel = torch.cat((start_state, torch.zeros(start_state.nelement(), **tensor_args)))
    cost_func_list.append(CostGP(robot, n_support_points, start_state_zero_vel, dt,
                                 dict(sigma_start=sigma_start, sigma_gp=sigma_gp), tensor_args=tensor_args))
    if multi_goal_states is not None:
        goal_zero_vel = torch.cat((multi_goal_states, torch.zeros_like(multi_goal_states)), dim=-1).unsqueeze(0)
        cost_func_list.append(CostGoalPrior(robot, n_support_points, multi_goal_states=goal_zero_vel.reshape(-1, goal_zero_vel.shape[-1]),
                                            num_particles_per_goal=num_particles_per_goal, num_samples=num_samples,
                                            sigma_goal_prior=sigma_goal_prior, tensor_args=tensor_args))
    for field in (collision_fields or []):
        cost_func_list.append(CostCollision(robot, n_support_points, field=field, sigma_coll=sigma_coll,
                                            tensor_args=tensor_args))
    if extra_costs:
        cost_func_list.extend(extra_costs)
    return CostComposite(robot, n_support_points, cost_func_list, tensor_args=tensor_args)
