"""Device-side collision field: one entry of ``collision_fields`` (mp_baselines/planners/gpmp2.py:72-79),
i.e. the object behind CostCollision / FieldFactor (costs/factors/field_factor.py:39), backed by the
primitive arrays of a ``mpb_field_desc``."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .models import ObstacleSet


class Field:
    """Anything a CostCollision can hold: produces a ``mpb_field_desc`` for the fused kernels."""

    robot = None

    def desc(self, weight=1.0, inv_sigma2=1.0):
        raise NotImplementedError

    def zero_grad(self):
        pass

    def bind_robot(self, robot):
        """The robot whose sphere table (radii, link indices) ``compute_cost`` refers to.  Set by the ``robot=``
        constructor argument or by the first CostCollision built around this field."""
        self.robot = robot
        return self

    def compute_cost(self, q_pos, link_pos, **kwargs):
        """The method ``FieldFactor.get_error`` calls (costs/factors/field_factor.py:39): hinge sum of every
        configuration, link_pos [..., Ns, ws_dim] -> [...].  Unknown keyword arguments are ignored (the reference
        forwards ``obstacle_spheres`` and, by accident, ``trajs_interp``: SURVEY quirk B6).  Differentiable w.r.t.
        ``link_pos`` (the reference takes the Jacobian by autograd, field_factor.py:52-57); the fused planners of this
        package never call it."""
        if self.robot is None:
            raise _lib.MpbError('compute_cost needs the robot sphere table: construct the field with robot=... '
                                '(or call field.bind_robot(robot))')
        return _FieldCost.apply(link_pos, self)


class _FieldCost(torch.autograd.Function):
    @staticmethod
    def forward(ctx, link_pos, field):
        robot = field.robot
        _lib.require_f32(link_pos)
        ns, ws = robot.model.n_spheres, robot.ws_dim
        if tuple(link_pos.shape[-2:]) != (ns, ws):
            raise _lib.MpbError(f'link_pos must end in [{ns}, {ws}] for robot {robot.name}, got {tuple(link_pos.shape)}')
        lp = link_pos.detach().reshape(-1, ns, ws).contiguous()
        err = torch.empty(lp.shape[0], device=lp.device, dtype=torch.float32)
        need = ctx.needs_input_grad[0]
        grad = torch.empty_like(lp) if need else None
        fdesc = field.desc()
        _lib.check(_lib.lib().mpb_field_cost(_lib.ptr(lp), lp.shape[0], robot.desc, fdesc, _lib.ptr(err), _lib.ptr(grad),
                                             _lib.stream_ptr()))
        ctx.grad_link = grad
        return err.view(link_pos.shape[:-2])

    @staticmethod
    def backward(ctx, grad_out):
        g = ctx.grad_link
        return (g * grad_out.reshape(-1, 1, 1)).view(*grad_out.shape, g.shape[-2], g.shape[-1]), None


def _cuda_args(tensor_args):
    if tensor_args is None:
        tensor_args = dict(device=torch.device('cuda', torch.cuda.current_device()), dtype=torch.float32)
    dev = torch.device(tensor_args['device'])
    if dev.type != 'cuda':
        raise _lib.MpbError('motion_planning_baselines_b200 fields live on a CUDA device (no CPU path)')
    return dict(device=dev, dtype=torch.float32)


class CollisionField(Field):
    def __init__(self, obstacles: ObstacleSet, tensor_args=None, robot=None):
        self.tensor_args = _cuda_args(tensor_args)
        self.robot = robot
        self.obstacles = obstacles
        self.cutoff_margin = obstacles.cutoff_margin
        ws = obstacles.ws_dim
        sph = np.zeros((obstacles.n_spheres, 4), np.float32)
        sph[:, :ws] = obstacles.sphere_centers
        sph[:, 3] = obstacles.sphere_radii
        box = np.zeros((obstacles.n_boxes, 8), np.float32)
        box[:, :ws] = obstacles.box_centers
        box[:, 4:4 + ws] = obstacles.box_half
        if ws == 2:
            box[:, 6] = np.inf          # the missing axis never constrains
        self._spheres = torch.tensor(sph, **self.tensor_args).contiguous()
        self._boxes = torch.tensor(box, **self.tensor_args).contiguous()

    def desc(self, weight=1.0, inv_sigma2=1.0):
        return _lib.FieldDesc(
            n_spheres=self.obstacles.n_spheres, n_boxes=self.obstacles.n_boxes,
            spheres=self._spheres.data_ptr() if self.obstacles.n_spheres else None,
            boxes=self._boxes.data_ptr() if self.obstacles.n_boxes else None,
            cutoff_margin=self.cutoff_margin, weight=weight, inv_sigma2=inv_sigma2, kind=_lib.FIELD_PRIMITIVES)


class SelfCollisionField(Field):
    """Self-collision of a sphere-model chain robot: sum over sphere pairs (i,j) on different links of
    relu(r_i + r_j + cutoff_margin - ||c_i - c_j||)  (the role of the external task's self-collision field,
    examples/panda_spheres_GPMP.py:41-45; MPB_FIELD_SELF).  ``pairs`` [Np,2] index the robot's sphere table;
    they are re-ordered here as the C ABI wants them: link(i) < link(j), sorted by (link(i), link(j))."""

    def __init__(self, robot_model, pairs=None, cutoff_margin=0.0, tensor_args=None, robot=None):
        self.tensor_args = _cuda_args(tensor_args)
        self.robot = robot
        if robot_model.kind != 'chain':
            raise _lib.MpbError('self-collision fields need a chain robot')
        if pairs is None:
            from .models import self_collision_pairs
            pairs = self_collision_pairs(robot_model)
        pairs = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
        link = np.asarray(robot_model.sphere_link)
        li, lj = link[pairs[:, 0]], link[pairs[:, 1]]
        if np.any(li == lj):
            raise _lib.MpbError('self-collision pairs must join spheres of different links')
        swap = li > lj
        pairs[swap] = pairs[swap][:, ::-1]
        li, lj = link[pairs[:, 0]], link[pairs[:, 1]]
        order = np.lexsort((pairs[:, 1], pairs[:, 0], lj, li))
        self.pairs = np.ascontiguousarray(pairs[order].astype(np.int32))
        if self.pairs.shape[0] > _lib.MPB_MAX_SELF_PAIRS:
            raise _lib.MpbError(f'at most {_lib.MPB_MAX_SELF_PAIRS} self-collision pairs')
        self.cutoff_margin = float(cutoff_margin)
        self._pairs = torch.tensor(self.pairs, device=self.tensor_args['device'], dtype=torch.int32).contiguous()

    def desc(self, weight=1.0, inv_sigma2=1.0):
        return _lib.FieldDesc(kind=_lib.FIELD_SELF, n_pairs=int(self.pairs.shape[0]),
                              pairs=self._pairs.data_ptr() if self.pairs.shape[0] else None,
                              cutoff_margin=self.cutoff_margin, weight=weight, inv_sigma2=inv_sigma2)


class WorkspaceBoundaryField(Field):
    """Workspace limits as a collision field: sdf(c) = min over axes of min(c - ws_min, ws_max - c), hinge
    relu(r_s + cutoff_margin - sdf) summed over the robot's spheres (the role of the external task's
    workspace-boundary field; MPB_FIELD_WORKSPACE)."""

    def __init__(self, ws_min, ws_max, cutoff_margin=0.0, tensor_args=None, robot=None):
        self.tensor_args = _cuda_args(tensor_args)
        self.robot = robot
        self.ws_min = [float(v) for v in np.asarray(ws_min).reshape(-1)]
        self.ws_max = [float(v) for v in np.asarray(ws_max).reshape(-1)]
        assert len(self.ws_min) == len(self.ws_max) and len(self.ws_min) in (2, 3)
        self.cutoff_margin = float(cutoff_margin)

    def desc(self, weight=1.0, inv_sigma2=1.0):
        lo = self.ws_min + [0.0] * (3 - len(self.ws_min))
        hi = self.ws_max + [1.0] * (3 - len(self.ws_max))
        return _lib.FieldDesc(kind=_lib.FIELD_WORKSPACE, ws_min=(C.c_float * 3)(*lo), ws_max=(C.c_float * 3)(*hi),
                              cutoff_margin=self.cutoff_margin, weight=weight, inv_sigma2=inv_sigma2)
