"""Batched collision queries and trajectory statistics: the slice of the external ``PlanningTask`` that the
reference planners and examples call -- ``task.compute_collision`` / ``random_coll_free_q`` / ``random_q`` /
``distance_q`` (mp_baselines/planners/rrt_base.py:57,101,107,110) and the statistics printed by every example
(examples/pointmass_dense_2d_CHOMP.py:130-133: fraction of free trajectories, collision intensity, success).
SURVEY 8f row 4.  The tree search of the RRT planners stays out of scope (sequential CPU code); what they need from
the GPU is this query: N configurations against every field in ONE launch (``mpb_collision_query``)."""
import ctypes as C

import torch

from . import _lib
from .fields import Field


class PlanningTask:
    def __init__(self, robot, collision_fields, ws_limits=None, tensor_args=None):
        self.robot = robot
        self.tensor_args = tensor_args if tensor_args is not None else robot.tensor_args
        self.fields = list(collision_fields)
        for f in self.fields:
            if not isinstance(f, Field):
                raise _lib.MpbError('PlanningTask needs motion_planning_baselines_b200.fields.Field objects')
            if f.robot is None:
                f.bind_robot(robot)
        if len(self.fields) > _lib.MPB_MAX_FIELDS:
            raise NotImplementedError(f'at most {_lib.MPB_MAX_FIELDS} collision fields')
        self.ws_limits = ws_limits

    def get_collision_fields(self):
        return self.fields

    def _descs(self):
        descs = [f.desc() for f in self.fields]
        return (_lib.FieldDesc * max(1, len(descs)))(*descs), len(descs)

    # ------------------------------------------------------------------ state queries
    def _query(self, q, row_stride, n, want_err=False):
        flag = torch.empty(n, device=q.device, dtype=torch.uint8)
        err = torch.empty(n, device=q.device, dtype=torch.float32) if want_err else None
        arr, nf = self._descs()
        _lib.check(_lib.lib().mpb_collision_query(_lib.ptr(q), n, row_stride, C.byref(self.robot.desc), arr, nf,
                                                  _lib.ptr(flag), _lib.ptr(err), _lib.stream_ptr()))
        return flag, err

    def compute_collision(self, q, **kwargs):
        """[..., >=d] joint states (positions first) -> bool [...]: True where any collision hinge is non-zero."""
        _lib.require_f32(q)
        qc = q.contiguous()
        stride = qc.shape[-1]
        n = qc.numel() // stride if stride else 0
        flag, _ = self._query(qc, stride, n)
        return flag.view(q.shape[:-1]).bool()

    def compute_collision_cost(self, q, **kwargs):
        """Sum over fields of the unweighted hinge sums of every state, [..., >=d] -> [...]."""
        _lib.require_f32(q)
        qc = q.contiguous()
        stride = qc.shape[-1]
        _, err = self._query(qc, stride, qc.numel() // stride, want_err=True)
        return err.view(q.shape[:-1])

    def random_q(self, n_samples=1, generator=None):
        lo, hi = self.robot.q_min, self.robot.q_max
        u = torch.rand(n_samples, self.robot.q_dim, generator=generator, **self.tensor_args)
        return lo + (hi - lo) * u

    def random_coll_free_q(self, n_samples=1, max_samples=1000, max_tries=1000, generator=None):
        """Rejection sampling in batches of ``max_samples`` (the contract of rrt_base.py:56-57)."""
        found, n_found = [], 0
        for _ in range(max_tries):
            qs = self.random_q(max_samples, generator=generator)
            free = qs[~self.compute_collision(qs)]
            found.append(free)
            n_found += free.shape[0]
            if n_found >= n_samples:
                break
        out = torch.cat(found, dim=0)[:n_samples]
        if out.shape[0] < n_samples:
            raise _lib.MpbError(f'found only {out.shape[0]} of {n_samples} collision-free configurations')
        return out.squeeze(0) if n_samples == 1 else out

    def distance_q(self, q1, q2):
        return torch.linalg.norm(q1 - q2, dim=-1)

    # ------------------------------------------------------------------ trajectory statistics
    def get_trajs_collision_and_free(self, trajs, return_indices=False):
        """trajs [B,H,>=d] -> (trajectories with at least one state in collision | None, free ones | None)."""
        coll = self.compute_collision(trajs)            # [B,H]
        any_coll = coll.any(dim=-1)
        idx_c, idx_f = torch.nonzero(any_coll).flatten(), torch.nonzero(~any_coll).flatten()
        tc = trajs[idx_c] if idx_c.numel() else None
        tf = trajs[idx_f] if idx_f.numel() else None
        if return_indices:
            return tc, idx_c, tf, idx_f, coll
        return tc, tf

    def compute_fraction_free_trajs(self, trajs):
        coll = self.compute_collision(trajs)
        return float((~coll.any(dim=-1)).float().mean())

    def compute_collision_intensity_trajs(self, trajs):
        """Fraction of waypoints in collision over the whole batch."""
        return float(self.compute_collision(trajs).float().mean())

    def compute_success_free_trajs(self, trajs):
        """1 if at least one trajectory of the batch is collision-free."""
        return int(bool((~self.compute_collision(trajs).any(dim=-1)).any()))
