"""Device-side robot objects: the duck-typed ``robot`` the reference planners and cost objects
take (mp_baselines/planners/costs/cost_functions.py:21,50-52,380,412-418), backed by a
``mpb_robot_desc`` for the fused kernels."""
import numpy as np
import torch

from . import _lib
from .models import RobotModel, panda_model, point_mass_model


class Robot:
    def __init__(self, model: RobotModel, dt=1.0, tensor_args=None):
        if tensor_args is None:
            tensor_args = dict(device=torch.device('cuda', torch.cuda.current_device()), dtype=torch.float32)
        dev = torch.device(tensor_args['device'])
        if dev.type != 'cuda':
            raise _lib.MpbError('motion_planning_baselines_b200 robots live on a CUDA device (no CPU path)')
        if tensor_args.get('dtype', torch.float32) != torch.float32:
            raise _lib.MpbError('the fused hot path computes in float32')
        self.model = model
        self.tensor_args = dict(device=dev, dtype=torch.float32)
        _lib.init_device(dev)
        self.q_dim = model.q_dim
        self.ws_dim = model.ws_dim
        self.dt = dt
        self.name = model.name
        self.q_min = torch.tensor(model.q_min, **self.tensor_args)
        self.q_max = torch.tensor(model.q_max, **self.tensor_args)
        self.link_radii = torch.tensor(model.sphere_r, **self.tensor_args)
        self._sphere_r = self.link_radii.contiguous()
        self._fixed_tf = self._sphere_link = self._sphere_off = None
        if model.kind == 'chain':
            self._fixed_tf = torch.tensor(model.fixed_tf, **self.tensor_args).contiguous()
            self._sphere_link = torch.tensor(model.sphere_link, device=dev, dtype=torch.int32).contiguous()
            self._sphere_off = torch.tensor(model.sphere_off, **self.tensor_args).contiguous()
            assert bool(np.all(np.diff(model.sphere_link) >= 0)), 'sphere table must be sorted by link'
        self.desc = _lib.RobotDesc(
            kind=_lib.ROBOT_CHAIN if model.kind == 'chain' else _lib.ROBOT_POINT,
            q_dim=model.q_dim, ws_dim=model.ws_dim, n_spheres=model.n_spheres,
            fixed_tf=None if self._fixed_tf is None else self._fixed_tf.data_ptr(),
            sphere_link=None if self._sphere_link is None else self._sphere_link.data_ptr(),
            sphere_off=None if self._sphere_off is None else self._sphere_off.data_ptr(),
            sphere_r=self._sphere_r.data_ptr())

    # --- the contract the reference's Cost classes use -------------------------------------
    def get_position(self, x):
        return x[..., :self.q_dim]

    def get_velocity(self, x):
        return x[..., self.q_dim:2 * self.q_dim]

    def fk_map_collision(self, q_pos, **kwargs):
        """Centres of the collision spheres, [..., d] -> [..., Ns, ws_dim]: the method the reference's
        ``Cost.get_q_pos_vel_and_fk_map`` calls (cost_functions.py:50-52).  The fused kernels never call it (they keep
        the link frames in registers); it exists so that an unmodified reference planner can use this robot.
        Differentiable: the backward is ``mpb_fk_spheres_vjp``."""
        if self.model.kind != 'chain':
            return q_pos[..., :self.ws_dim].unsqueeze(-2)
        return _FkSpheres.apply(q_pos, self)


class _FkSpheres(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q_pos, robot):
        _lib.require_f32(q_pos)
        q = q_pos.detach().reshape(-1, robot.q_dim).contiguous()
        out = torch.empty(q.shape[0], robot.model.n_spheres, 3, device=q.device, dtype=torch.float32)
        _lib.check(_lib.lib().mpb_fk_spheres(_lib.ptr(q), q.shape[0], robot.desc, _lib.ptr(out), _lib.stream_ptr()))
        ctx.robot, ctx.q = robot, q
        return out.view(*q_pos.shape[:-1], robot.model.n_spheres, 3)

    @staticmethod
    def backward(ctx, grad_out):
        robot, q = ctx.robot, ctx.q
        g = grad_out.reshape(-1, robot.model.n_spheres, 3).contiguous().float()
        gq = torch.empty_like(q)
        _lib.check(_lib.lib().mpb_fk_spheres_vjp(_lib.ptr(q), _lib.ptr(g), q.shape[0], robot.desc, _lib.ptr(gq), _lib.stream_ptr()))
        return gq.view(*grad_out.shape[:-2], robot.q_dim), None


def RobotPointMass(q_dim=2, radius=0.0, dt=1.0, tensor_args=None):
    return Robot(point_mass_model(q_dim, radius), dt=dt, tensor_args=tensor_args)


def RobotPanda(dt=1.0, tensor_args=None):
    return Robot(panda_model(), dt=dt, tensor_args=tensor_args)
