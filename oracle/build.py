"""Build oracle robot / field objects from the product's plain parameter holders
(TEST INFRASTRUCTURE -- see oracle/__init__.py)."""
import torch

from .fields import PrimitiveField, SelfCollisionField, WorkspaceBoundaryField
from .robots import PointMassRobot, SerialChainRobot

TA = dict(device='cpu', dtype=torch.float32)


def oracle_robot(model, dt, tensor_args=TA):
    if model.kind == 'point':
        return PointMassRobot(model.q_dim, radius=float(model.sphere_r[0]), dt=dt, tensor_args=tensor_args)
    return SerialChainRobot(model.fixed_tf, model.sphere_link, model.sphere_off, model.sphere_r,
                            model.q_min, model.q_max, dt=dt, tensor_args=tensor_args)


def oracle_field(obst, model, tensor_args=TA):
    return PrimitiveField(sphere_centers=obst.sphere_centers, sphere_radii=obst.sphere_radii,
                          box_centers=obst.box_centers, box_half=obst.box_half,
                          link_radii=model.sphere_r, cutoff_margin=obst.cutoff_margin,
                          ws_dim=obst.ws_dim, tensor_args=tensor_args)


def oracle_self_field(pairs, model, cutoff_margin, tensor_args=TA):
    return SelfCollisionField(pairs, model.sphere_r, cutoff_margin=cutoff_margin, tensor_args=tensor_args)


def oracle_workspace_field(ws_min, ws_max, model, cutoff_margin, tensor_args=TA):
    return WorkspaceBoundaryField(ws_min, ws_max, link_radii=model.sphere_r, cutoff_margin=cutoff_margin,
                                  tensor_args=tensor_args)
