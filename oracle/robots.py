"""Oracle robots (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Duck-typed to what the reference planners touch on a ``robot``
(mp_baselines/planners/costs/cost_functions.py:21,50-52 -- ``q_dim``,
``get_position``, ``get_velocity``, ``fk_map_collision``; cost_functions.py:380
``dt``; cost_functions.py:412-418 ``q_min``/``q_max``).

PARITY UNPINNED: the reference's FK lives in the absent ``torch_robotics``
dependency; these classes are the specification.  All maths is plain eager
torch so that it is differentiable (GPMP2 / CHOMP take autograd gradients
through it, field_factor.py:54, chomp.py:139).
"""
import torch


class PointMassRobot:
    """q is the workspace position: one collision sphere at q."""

    def __init__(self, q_dim, radius=0.0, dt=1.0, q_limit=1.0, tensor_args=None):
        self.q_dim = q_dim
        self.ws_dim = q_dim
        self.dt = dt
        self.tensor_args = tensor_args or dict(device='cpu', dtype=torch.float32)
        self.q_min = torch.full((q_dim,), -q_limit, **self.tensor_args)
        self.q_max = torch.full((q_dim,), q_limit, **self.tensor_args)
        self.link_radii = torch.full((1,), radius, **self.tensor_args)

    def get_position(self, x):
        return x[..., :self.q_dim]

    def get_velocity(self, x):
        return x[..., self.q_dim:2 * self.q_dim]

    def fk_map_collision(self, q_pos):
        return q_pos.unsqueeze(-2)


class SerialChainRobot:
    """Serial chain of revolute-z joints (URDF convention) with collision spheres.

    Frame of link j:  T_j = T_{j-1} @ F_j @ Rz(q_j)  with F_j the fixed 3x4
    parent->joint transform.  Sphere s sits at ``T_{link[s]} @ [offset_s, 1]``.

    fixed_tf     [J,3,4] float   fixed parent->joint transforms
    sphere_link  [Ns]    int     0-based index of the joint whose frame carries the sphere
    sphere_off   [Ns,3]  float   centre in that frame
    sphere_r     [Ns]    float   radius
    """

    def __init__(self, fixed_tf, sphere_link, sphere_off, sphere_r, q_min, q_max, dt=1.0, tensor_args=None):
        self.tensor_args = tensor_args or dict(device='cpu', dtype=torch.float32)
        self.fixed_tf = torch.as_tensor(fixed_tf).to(**self.tensor_args)
        self.sphere_link = torch.as_tensor(sphere_link).to(device=self.tensor_args['device'], dtype=torch.int64)
        self.sphere_off = torch.as_tensor(sphere_off).to(**self.tensor_args)
        self.link_radii = torch.as_tensor(sphere_r).to(**self.tensor_args)
        self.q_dim = self.fixed_tf.shape[0]
        self.ws_dim = 3
        self.dt = dt
        self.q_min = torch.as_tensor(q_min).to(**self.tensor_args)
        self.q_max = torch.as_tensor(q_max).to(**self.tensor_args)

    def get_position(self, x):
        return x[..., :self.q_dim]

    def get_velocity(self, x):
        return x[..., self.q_dim:2 * self.q_dim]

    def link_frames(self, q_pos):
        """Rotation [..., J, 3, 3] and origin [..., J, 3] of every link frame."""
        R = torch.eye(3, **self.tensor_args).expand(*q_pos.shape[:-1], 3, 3)
        t = torch.zeros(*q_pos.shape[:-1], 3, **self.tensor_args)
        Rs, ts = [], []
        c_all, s_all = torch.cos(q_pos), torch.sin(q_pos)
        for j in range(self.q_dim):
            Fr, Ft = self.fixed_tf[j, :, :3], self.fixed_tf[j, :, 3]
            t = t + (R @ Ft.unsqueeze(-1)).squeeze(-1)
            R = R @ Fr
            c, s = c_all[..., j, None], s_all[..., j, None]
            col0 = R[..., :, 0] * c + R[..., :, 1] * s
            col1 = R[..., :, 1] * c - R[..., :, 0] * s
            R = torch.stack((col0, col1, R[..., :, 2]), dim=-1)
            Rs.append(R)
            ts.append(t)
        return torch.stack(Rs, dim=-3), torch.stack(ts, dim=-2)

    def fk_map_collision(self, q_pos):
        R, t = self.link_frames(q_pos)                       # [...,J,3,3], [...,J,3]
        Rs = R.index_select(-3, self.sphere_link)            # [...,Ns,3,3]
        ts = t.index_select(-2, self.sphere_link)            # [...,Ns,3]
        return (Rs @ self.sphere_off.unsqueeze(-1)).squeeze(-1) + ts
