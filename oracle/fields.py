"""Oracle collision field (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Duck-typed to ``field.compute_cost(q_pos, link_pos, **kw)`` / ``field.zero_grad()``
(mp_baselines/planners/costs/factors/field_factor.py:39,56).

PARITY UNPINNED: the reference's MultiSphereField / MultiBoxField and the
collision hinge live in the absent ``torch_robotics`` dependency; this file is
the specification (SURVEY.md Appendix C semantics):

    sphere_sdf_i(x) = ||x - c_i|| - r_i
    box_sdf_j(x)    = || max(q,0) || + min(max_k q_k, 0),   q = |x - c_j| - half_j
    sdf(x)          = min over all primitives
    cost[b,t]       = sum_s relu( (radius_s + cutoff_margin) - sdf(link_pos[b,t,s]) )

The fp32 evaluation ORDER below is part of the spec: every operation is a
separately rounded eager op (square roots included, see _sqrt), squares are summed left to right.  The CUDA
kernel evaluates potentially colliding pairs with the same correctly rounded
operations in the same order, which is what makes hinge terms and
collision-free flags bit-identical for point-mass robots.
"""
import torch


# torch.sqrt on CPU float32 tensors is NOT correctly rounded in this image (the vectorised body of the loop is off by
# one ulp for ~0.7 % of random inputs, the scalar tail is exact -- so results even depend on how the work is split
# over threads), whereas CUDA's sqrtf -- what the reference runs on a GPU -- is IEEE-754 correctly rounded.  The spec
# is the IEEE result: sqrt in float64, rounded once to float32 (double rounding is harmless for sqrt at 53 >= 2*24+2
# bits).  bench.py's CPU-baseline timing switches this off to keep the reference's own cost model.
EXACT_SQRT = True


def _sqrt(v):
    if EXACT_SQRT and v.dtype == torch.float32:
        return torch.sqrt(v.double()).to(torch.float32)
    return torch.sqrt(v)


def _sum_sq(v):
    acc = v[..., 0] * v[..., 0]
    for k in range(1, v.shape[-1]):
        acc = acc + v[..., k] * v[..., k]
    return acc


class PrimitiveField:
    def __init__(self, sphere_centers=None, sphere_radii=None, box_centers=None, box_half=None,
                 link_radii=None, cutoff_margin=0.0, ws_dim=None, tensor_args=None):
        self.tensor_args = tensor_args or dict(device='cpu', dtype=torch.float32)
        ta = self.tensor_args

        def _t(a, cols):
            if a is None:
                return torch.zeros((0,) + cols, **ta)
            return torch.as_tensor(a).to(**ta).reshape((-1,) + cols)

        if ws_dim is None:
            src = sphere_centers if sphere_centers is not None else box_centers
            ws_dim = torch.as_tensor(src).reshape(len(src), -1).shape[-1]
        self.ws_dim = ws_dim
        self.sphere_centers = _t(sphere_centers, (ws_dim,))
        self.sphere_radii = _t(sphere_radii, ())
        self.box_centers = _t(box_centers, (ws_dim,))
        self.box_half = _t(box_half, (ws_dim,))
        self.link_radii = torch.as_tensor(link_radii if link_radii is not None else [0.0]).to(**ta).reshape(-1)
        self.cutoff_margin = float(cutoff_margin)

    # ------------------------------------------------------------------ sdf
    def sdf(self, x):
        """x [..., ws] -> [...] signed distance to the union of primitives."""
        out = None
        if self.sphere_centers.shape[0]:
            d = x.unsqueeze(-2) - self.sphere_centers           # [..., Nsph, ws]
            s = _sqrt(_sum_sq(d)) - self.sphere_radii
            out = s.min(dim=-1).values
        if self.box_centers.shape[0]:
            q = (x.unsqueeze(-2) - self.box_centers).abs() - self.box_half
            s2 = _sum_sq(torch.clamp(q, min=0.0))
            # same forward values as sqrt(s2); where s2 == 0 (point inside the box) no gradient flows through the
            # square root (plain sqrt would produce 0 * inf = NaN under autograd), leaving d/dx max_k q_k
            pos = s2 > 0
            outside = _sqrt(torch.where(pos, s2, torch.ones_like(s2))) * pos
            inside = torch.clamp(q.max(dim=-1).values, max=0.0)
            b = (outside + inside).min(dim=-1).values
            out = b if out is None else torch.minimum(out, b)
        if out is None:
            out = torch.full(x.shape[:-1], float('inf'), **self.tensor_args)
        return out

    def hinge_terms(self, link_pos):
        """link_pos [..., Ns, ws] -> per-sphere hinge [..., Ns]."""
        thr = self.link_radii + torch.tensor(self.cutoff_margin, **self.tensor_args)
        return torch.relu(thr - self.sdf(link_pos))

    def compute_cost(self, q_pos, link_pos, **kwargs):
        """-> [B, H'] (sum over the robot's collision spheres). Unknown kwargs are
        accepted and ignored (reference quirk B6: ``trajs_interp`` leaks in here)."""
        h = self.hinge_terms(link_pos)
        acc = h[..., 0]
        for s in range(1, h.shape[-1]):
            acc = acc + h[..., s]
        return acc

    def collision_free(self, link_pos):
        """link_pos [B, H', Ns, ws] -> bool [B]: every hinge term is exactly zero."""
        return (self.hinge_terms(link_pos) == 0).flatten(1).all(dim=1)

    def zero_grad(self):
        pass


class SelfCollisionField:
    """Self-collision of a sphere-model robot (the role of the external task's self-collision field,
    examples/panda_spheres_GPMP.py:41-45 ``use_self_collision_storm``; PARITY UNPINNED, this is the spec):

        cost[b,t] = sum over pairs (i,j) of relu( (r_i + r_j + cutoff_margin) - ||c_i - c_j|| )

    pairs [Np,2] int: indices into the robot's collision-sphere table, spheres on different links."""

    def __init__(self, pairs, link_radii, cutoff_margin=0.0, tensor_args=None):
        self.tensor_args = tensor_args or dict(device='cpu', dtype=torch.float32)
        self.pairs = torch.as_tensor(pairs).to(device=self.tensor_args['device'], dtype=torch.int64).reshape(-1, 2)
        self.link_radii = torch.as_tensor(link_radii).to(**self.tensor_args).reshape(-1)
        self.cutoff_margin = float(cutoff_margin)

    def hinge_terms(self, link_pos):
        """link_pos [..., Ns, 3] -> per-pair hinge [..., Np]."""
        i, j = self.pairs[:, 0], self.pairs[:, 1]
        d = link_pos.index_select(-2, i) - link_pos.index_select(-2, j)
        dist = _sqrt(_sum_sq(d))
        thr = (self.link_radii[i] + self.link_radii[j]) + torch.tensor(self.cutoff_margin, **self.tensor_args)
        return torch.relu(thr - dist)

    def compute_cost(self, q_pos, link_pos, **kwargs):
        return self.hinge_terms(link_pos).sum(-1)

    def collision_free(self, link_pos):
        return (self.hinge_terms(link_pos) == 0).flatten(1).all(dim=1)

    def zero_grad(self):
        pass


class WorkspaceBoundaryField:
    """Workspace limits as a collision field (the role of the external task's workspace-boundary field; PARITY
    UNPINNED, this is the spec):

        sdf(c)    = min( c_x - lo_x, c_y - lo_y, [c_z - lo_z,] hi_x - c_x, hi_y - c_y [, hi_z - c_z] )
        cost[b,t] = sum_s relu( (radius_s + cutoff_margin) - sdf(link_pos[b,t,s]) )

    torch.min returns the FIRST minimal entry, which fixes the sub-gradient at ties."""

    def __init__(self, ws_min, ws_max, link_radii=None, cutoff_margin=0.0, tensor_args=None):
        self.tensor_args = tensor_args or dict(device='cpu', dtype=torch.float32)
        self.ws_min = torch.as_tensor(ws_min).to(**self.tensor_args).reshape(-1)
        self.ws_max = torch.as_tensor(ws_max).to(**self.tensor_args).reshape(-1)
        self.ws_dim = self.ws_min.shape[0]
        self.link_radii = torch.as_tensor(link_radii if link_radii is not None else [0.0]).to(**self.tensor_args).reshape(-1)
        self.cutoff_margin = float(cutoff_margin)

    def sdf(self, x):
        return torch.cat((x - self.ws_min, self.ws_max - x), dim=-1).min(dim=-1).values

    def hinge_terms(self, link_pos):
        thr = self.link_radii + torch.tensor(self.cutoff_margin, **self.tensor_args)
        return torch.relu(thr - self.sdf(link_pos))

    def compute_cost(self, q_pos, link_pos, **kwargs):
        h = self.hinge_terms(link_pos)
        acc = h[..., 0]
        for s in range(1, h.shape[-1]):
            acc = acc + h[..., s]
        return acc

    def collision_free(self, link_pos):
        return (self.hinge_terms(link_pos) == 0).flatten(1).all(dim=1)

    def zero_grad(self):
        pass
