"""Oracle cost terms (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Functional restatement of
  * CostGP.eval / get_linear_system          mp_baselines/planners/costs/cost_functions.py:271-314
  * CostGoalPrior.eval / get_linear_system   cost_functions.py:523-554
  * CostCollision.eval / get_linear_system   cost_functions.py:171-231 (+ field_factor.py:17-57)
  * CostComposite.eval / get_linear_system   cost_functions.py:70-144
  * build_gpmp2_cost_composite defaults      mp_baselines/planners/gpmp2.py:23-89
  * CostGPTrajectory / CostSmoothnessCHOMP / CostJointLimits .eval   cost_functions.py:317-429
Pinned by tests/golden/*.npz (generated from the unmodified reference).
"""
import torch

from .gp_prior import gp_Q_inv, phi_matrix, unary_K


def interpolate_points(trajs, n_interpolated_points):
    """Linear joint-space up-sampling with n extra points between consecutive waypoints: [B,H,D] ->
    [B,(H-1)(n+1)+1,D].  The reference imports ``interpolate_points_v1`` from the absent torch_robotics
    (cost_functions.py:13,118): PARITY UNPINNED, this is the spec (same as oracle/ref_shim)."""
    B, H, D = trajs.shape
    w = torch.linspace(0, 1, n_interpolated_points + 2, dtype=trajs.dtype, device=trajs.device)[:-1]
    seg = trajs[:, :-1].unsqueeze(2) * (1 - w).view(1, 1, -1, 1) + trajs[:, 1:].unsqueeze(2) * w.view(1, 1, -1, 1)
    return torch.cat((seg.reshape(B, -1, D), trajs[:, -1:]), dim=1)


def joint_limits_cost(x, q_min, q_max, eps):
    """CostJointLimits.eval (cost_functions.py:393-426): squared violation of the limits shrunk by eps, SUMMED OVER
    THE WHOLE BATCH -- the reference's ``.sum(-1)`` runs over the flat list of violating entries, so the result
    is a 0-dim tensor that a composite adds to every trajectory (quirk B14)."""
    d = q_min.shape[0]
    q = x[..., :d]
    lower = torch.relu((q_min + eps) - q)
    upper = torch.relu(q - (q_max - eps))
    return (lower ** 2).sum() + (upper ** 2).sum()


def smoothness_chomp_cost(x, R):
    """CostSmoothnessCHOMP.eval (cost_functions.py:371-387): x[:, :, j]^T R x[:, :, j] per trajectory and state
    column -> [B, D] (``batched_weighted_dot_prod`` is external; semantics as in oracle/ref_shim)."""
    return (x.transpose(-1, -2) @ R.unsqueeze(0) @ x).diagonal(dim1=-2, dim2=-1)


def gp_trajectory_cost(x, Phi, Q_inv):
    """CostGPTrajectory.eval (cost_functions.py:343-354): the GP-prior part of CostGP without the start term."""
    D = x.shape[-1]
    e = (x[:, 1:].unsqueeze(-1) - Phi @ x[:, :-1].unsqueeze(-1))
    return (e.transpose(2, 3) @ Q_inv.reshape(1, 1, D, D) @ e).sum(1).squeeze()


class CostSpec:
    """The GPMP2 / Stoch-GPMP composite: start prior + GP prior + goal prior + one collision
    term per field, all with weight 1 (gpmp2.py:23-89)."""

    def __init__(self, robot, H, dt, start_state, goal_state, fields,
                 sigma_start=1e-5, sigma_gp=1e-2, sigma_coll=1e-5, sigma_goal_prior=1e-5, tensor_args=None):
        ta = tensor_args
        d = robot.q_dim
        self.robot, self.H, self.d, self.D, self.dt, self.tensor_args = robot, H, d, 2 * d, dt, ta
        self.fields = list(fields)
        zeros = torch.zeros(d, **ta)
        self.start_zero_vel = torch.cat((start_state[:d].to(**ta), zeros))
        self.goal_zero_vel = None if goal_state is None else torch.cat((goal_state[:d].to(**ta), zeros))
        self.K_start = unary_K(self.D, sigma_start, ta)
        self.K_goal = unary_K(self.D, sigma_goal_prior, ta)
        self.Q_inv = gp_Q_inv(d, dt, sigma_gp, ta)
        self.Phi = phi_matrix(d, dt, ta)
        self.w_coll = 1. / (sigma_coll ** 2)

    # ------------------------------------------------------------- eval
    def gp_cost(self, x):
        e0 = (self.start_zero_vel - x[:, 0]).unsqueeze(1)                      # [B,1,D]
        start = (e0 @ self.K_start.unsqueeze(0) @ e0.transpose(1, 2)).reshape(-1)
        e = (x[:, 1:].unsqueeze(-1) - self.Phi @ x[:, :-1].unsqueeze(-1))      # [B,H-1,D,1]
        gp = (e.transpose(2, 3) @ self.Q_inv.reshape(1, 1, self.D, self.D) @ e).sum(1).reshape(-1)
        return start + gp

    def goal_cost(self, x):
        eg = (self.goal_zero_vel - x[:, -1]).unsqueeze(1)
        return (eg @ self.K_goal.unsqueeze(0) @ eg.transpose(1, 2)).reshape(-1)

    def collision_errors(self, x, field):
        """[B, H-1]: waypoint 0 is skipped (cost_functions.py:165-169)."""
        q = self.robot.get_position(x)
        link_pos = self.robot.fk_map_collision(q)
        return field.compute_cost(q[:, 1:], link_pos[:, 1:]).reshape(x.shape[0], self.H - 1)

    def collision_cost(self, x, field):
        return self.w_coll * self.collision_errors(x, field).sum(1)

    def terms(self, x):
        out = [self.gp_cost(x)]
        if self.goal_zero_vel is not None:
            out.append(self.goal_cost(x))
        for f in self.fields:
            out.append(self.collision_cost(x, f))
        return out

    def eval(self, x):
        """x [B,H,D] or [N,B,H,D] -> [B] (or [N*B]); terms added left to right starting from 0."""
        if x.ndim == 4:
            x = x.reshape(-1, *x.shape[2:])
        total = 0
        for c in self.terms(x):
            total = total + 1.0 * c
        return total

    __call__ = eval

    def collision_free(self, x):
        if x.ndim == 4:
            x = x.reshape(-1, *x.shape[2:])
        q = self.robot.get_position(x)
        link_pos = self.robot.fk_map_collision(q)[:, 1:]
        free = torch.ones(x.shape[0], dtype=torch.bool)
        for f in self.fields:
            free &= f.collision_free(link_pos)
        return free

    # ------------------------------------------------- linear system (GPMP2)
    def linear_system(self, x, n_interpolated_points=None):
        """Dense (A, b, K) with rows [start D | GP (H-1)D | goal D | (H-1) per field]
        (cost_functions.py:107-144, 291-314, 538-554, 191-231).  With n_interpolated_points the collision
        Jacobian is that of the summed error over the linearly up-sampled trajectory (first point dropped) w.r.t.
        the support points, while b stays the error AT the support points (field_factor.py:41-57)."""
        ta, D, H, d = self.tensor_args, self.D, self.H, self.d
        B, N = x.shape[0], self.D * self.H
        x = x.detach().clone().requires_grad_(True)
        As, bs, Ks = [], [], []

        A = torch.zeros(B, N, N, **ta)
        b = torch.zeros(B, N, 1, **ta)
        K = torch.zeros(B, N, N, **ta)
        eye = torch.eye(D, **ta)
        A[:, :D, :D] = eye
        b[:, :D, 0] = self.start_zero_vel - x[:, 0]
        K[:, :D, :D] = self.K_start
        e = (x[:, 1:].unsqueeze(-1) - self.Phi @ x[:, :-1].unsqueeze(-1))
        for t in range(H - 1):
            r = slice(D * (t + 1), D * (t + 2))
            A[:, r, D * t:D * (t + 1)] = self.Phi
            A[:, r, r] = -eye
            K[:, r, r] = self.Q_inv
        b[:, D:, 0] = e.reshape(B, -1)
        As.append(A), bs.append(b), Ks.append(K)

        if self.goal_zero_vel is not None:
            A = torch.zeros(B, D, N, **ta)
            A[:, :, -D:] = eye
            b = (self.goal_zero_vel - x[:, -1]).reshape(B, D, 1)
            K = self.K_goal.expand(B, D, D)
            As.append(A), bs.append(b), Ks.append(K)

        x_interp = None if n_interpolated_points is None else interpolate_points(x, n_interpolated_points)
        for f in self.fields:
            err = self.collision_errors(x, f)                                 # [B,H-1]
            if x_interp is None:
                err_j = err
            else:
                qi = self.robot.get_position(x_interp)
                li = self.robot.fk_map_collision(qi)
                err_j = f.compute_cost(qi[:, 1:], li[:, 1:], trajs_interp=x_interp).reshape(B, -1)   # quirk B6: the kwarg leaks
            grad = torch.autograd.grad(err_j.sum(), x, retain_graph=True)[0]
            Hobs = -grad[:, 1:, :d]                                           # [B,H-1,d]
            A = torch.zeros(B, H - 1, N, **ta)
            for t in range(H - 1):
                A[:, t, D * (t + 1):D * (t + 1) + d] = Hobs[:, t]
            K = self.w_coll * torch.eye(H - 1, **ta).expand(B, H - 1, H - 1)
            As.append(A), bs.append(err.detach().unsqueeze(-1)), Ks.append(K)

        A = torch.cat([a.detach() for a in As], dim=1)
        b = torch.cat([v.detach() for v in bs], dim=1)
        R = A.shape[1]
        K = torch.zeros(B, R, R, **ta)
        o = 0
        for k in Ks:
            n = k.shape[1]
            K[:, o:o + n, o:o + n] = k
            o += n
        return A, b, K
