"""Device-side collision field: one entry of ``collision_fields`` (mp_baselines/planners/gpmp2.py:72-79),
i.e. the object behind CostCollision / FieldFactor (costs/factors/field_factor.py:39), backed by the
primitive arrays of a ``mpb_field_desc``."""
import numpy as np
import torch

from . import _lib
from .models import ObstacleSet


class CollisionField:
    def __init__(self, obstacles: ObstacleSet, tensor_args=None):
        if tensor_args is None:
            tensor_args = dict(device=torch.device('cuda', torch.cuda.current_device()), dtype=torch.float32)
        dev = torch.device(tensor_args['device'])
        if dev.type != 'cuda':
            raise _lib.MpbError('motion_planning_baselines_b200 fields live on a CUDA device (no CPU path)')
        self.obstacles = obstacles
        self.tensor_args = dict(device=dev, dtype=torch.float32)
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
            cutoff_margin=self.cutoff_margin, weight=weight, inv_sigma2=inv_sigma2)

    def zero_grad(self):
        pass
