"""Robot kinematics + collision-sphere tables and obstacle-set descriptors.

These are plain parameter holders (numpy on the host); the CUDA side receives them through
``mpb_robot_desc`` / ``mpb_field_desc`` (include/mpb.h).  They play the role of the
``torch_robotics`` robot / environment objects the reference examples construct
(e.g. examples/panda_spheres_GPMP.py:31-57), which are NOT part of the reference repo:
the Panda chain below is the public franka_description URDF and the
collision-sphere table is our own deterministic 50-sphere model.
"""
import math

import numpy as np


def _rpy_xyz_to_tf(xyz, rpy):
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    R = np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                  [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                  [-sp, cp * sr, cp * cr]], dtype=np.float64)
    R[np.abs(R) < 1e-15] = 0.0          # rpy are multiples of pi/2: make the zeros exact
    T = np.zeros((3, 4), dtype=np.float64)
    T[:, :3] = R
    T[:, 3] = xyz
    return T


# (xyz, rpy) of the seven revolute-z joints, parent link frame -> joint frame.
PANDA_JOINT_ORIGINS = [
    ((0.0, 0.0, 0.333), (0.0, 0.0, 0.0)),
    ((0.0, 0.0, 0.0), (-math.pi / 2, 0.0, 0.0)),
    ((0.0, -0.316, 0.0), (math.pi / 2, 0.0, 0.0)),
    ((0.0825, 0.0, 0.0), (math.pi / 2, 0.0, 0.0)),
    ((-0.0825, 0.384, 0.0), (-math.pi / 2, 0.0, 0.0)),
    ((0.0, 0.0, 0.0), (math.pi / 2, 0.0, 0.0)),
    ((0.088, 0.0, 0.0), (math.pi / 2, 0.0, 0.0)),
]
PANDA_Q_MIN = np.array([-2.8973, -1.7628, -2.8973, -3.0718, -2.8973, -0.0175, -2.8973], dtype=np.float32)
PANDA_Q_MAX = np.array([2.8973, 1.7628, 2.8973, -0.0698, 2.8973, 3.7525, 2.8973], dtype=np.float32)

# Collision model: per link a segment a->b (link frame) covered by n equally spaced spheres.
# The hand group is given in the hand frame (flange +0.107 z, then -pi/4 about z) and is
# re-expressed in the link-7 frame so that the chain has exactly seven frames.
_PANDA_SEGMENTS = [
    # link, a, b, n, radius
    (0, (0.0, 0.0, -0.20), (0.0, 0.0, -0.02), 5, 0.075),
    (1, (0.0, -0.02, 0.0), (0.0, -0.20, 0.0), 6, 0.075),
    (2, (0.0, 0.0, -0.12), (0.08, 0.03, 0.0), 5, 0.070),
    (3, (0.0, 0.0, 0.0), (-0.0825, 0.12, 0.0), 4, 0.070),
    (4, (0.0, 0.02, -0.30), (0.0, 0.0, 0.0), 9, 0.065),
    (5, (0.0, 0.0, 0.0), (0.088, 0.0, 0.0), 3, 0.060),
    (6, (0.0, 0.0, 0.0), (0.0, 0.0, 0.107), 5, 0.055),
]
_PANDA_HAND_SEGMENTS = [
    ((0.0, -0.08, 0.03), (0.0, 0.08, 0.03), 5, 0.040),      # palm
    ((0.0, -0.04, 0.07), (0.0, -0.04, 0.12), 4, 0.022),     # finger
    ((0.0, 0.04, 0.07), (0.0, 0.04, 0.12), 4, 0.022),       # finger
]


class RobotModel:
    """kind 'point': q IS the workspace position (one sphere of radius ``radius``).
    kind 'chain': serial revolute-z chain with a sphere table."""

    def __init__(self, kind, q_dim, ws_dim, fixed_tf=None, sphere_link=None, sphere_off=None,
                 sphere_r=None, q_min=None, q_max=None, name=''):
        self.kind, self.q_dim, self.ws_dim, self.name = kind, q_dim, ws_dim, name
        self.fixed_tf = None if fixed_tf is None else np.ascontiguousarray(fixed_tf, dtype=np.float32)
        self.sphere_link = None if sphere_link is None else np.ascontiguousarray(sphere_link, dtype=np.int32)
        self.sphere_off = None if sphere_off is None else np.ascontiguousarray(sphere_off, dtype=np.float32)
        self.sphere_r = np.ascontiguousarray(sphere_r, dtype=np.float32)
        self.q_min = np.ascontiguousarray(q_min, dtype=np.float32)
        self.q_max = np.ascontiguousarray(q_max, dtype=np.float32)

    @property
    def n_spheres(self):
        return int(self.sphere_r.shape[0])


def point_mass_model(q_dim, radius=0.0, q_limit=1.0):
    assert q_dim in (2, 3)
    return RobotModel('point', q_dim, q_dim, sphere_r=[radius],
                      q_min=[-q_limit] * q_dim, q_max=[q_limit] * q_dim, name=f'pointmass{q_dim}d')


def panda_model():
    tfs = np.stack([_rpy_xyz_to_tf(xyz, rpy) for xyz, rpy in PANDA_JOINT_ORIGINS])
    link, off, rad = [], [], []
    for l, a, b, n, r in _PANDA_SEGMENTS:
        for k in range(n):
            s = k / (n - 1)
            link.append(l)
            off.append([a[i] + s * (b[i] - a[i]) for i in range(3)])
            rad.append(r)
    hand = _rpy_xyz_to_tf((0.0, 0.0, 0.107), (0.0, 0.0, -math.pi / 4))
    for a, b, n, r in _PANDA_HAND_SEGMENTS:
        for k in range(n):
            s = k / (n - 1)
            p = np.array([a[i] + s * (b[i] - a[i]) for i in range(3)])
            link.append(6)
            off.append(list(hand[:, :3] @ p + hand[:, 3]))
            rad.append(r)
    order = np.argsort(np.array(link), kind='stable')
    return RobotModel('chain', 7, 3, fixed_tf=tfs, sphere_link=np.array(link)[order],
                      sphere_off=np.array(off)[order], sphere_r=np.array(rad)[order],
                      q_min=PANDA_Q_MIN, q_max=PANDA_Q_MAX, name='panda')


# Link pairs checked for self-collision.  Like an SRDF's disabled pairs: adjacent links are skipped, and so are
# (1,3), (4,6) (their sphere groups overlap in every configuration) and (2,4) (overlap in 79 % of uniformly drawn
# configurations: the elbow) -- measured with this sphere table over 20 000 random joint vectors.
PANDA_SELF_LINK_PAIRS = [(0, 2), (0, 3), (0, 4), (0, 5), (0, 6), (1, 4), (1, 5), (1, 6), (2, 5), (2, 6), (3, 5), (3, 6)]


def self_collision_pairs(model, link_pairs=None):
    """[Np,2] sphere-index pairs (i on link a, j on link b) for every allowed link pair (a,b), sorted by link pair."""
    if link_pairs is None:
        link_pairs = PANDA_SELF_LINK_PAIRS if model.name == 'panda' else \
            [(a, b) for a in range(model.q_dim) for b in range(a + 2, model.q_dim)]
    link = np.asarray(model.sphere_link)
    out = []
    for a, b in sorted(link_pairs):
        for i in np.nonzero(link == a)[0]:
            for j in np.nonzero(link == b)[0]:
                out.append((int(i), int(j)))
    return np.asarray(out, dtype=np.int32).reshape(-1, 2)


class ObstacleSet:
    """Union of sphere and axis-aligned box primitives in a ws_dim workspace
    (the role of MultiSphereField / MultiBoxField behind CostCollision)."""

    def __init__(self, ws_dim, sphere_centers=None, sphere_radii=None, box_centers=None, box_half=None,
                 cutoff_margin=0.0, name=''):
        self.ws_dim = ws_dim
        self.sphere_centers = np.zeros((0, ws_dim), np.float32) if sphere_centers is None else \
            np.ascontiguousarray(sphere_centers, dtype=np.float32).reshape(-1, ws_dim)
        self.sphere_radii = np.zeros((0,), np.float32) if sphere_radii is None else \
            np.ascontiguousarray(sphere_radii, dtype=np.float32).reshape(-1)
        self.box_centers = np.zeros((0, ws_dim), np.float32) if box_centers is None else \
            np.ascontiguousarray(box_centers, dtype=np.float32).reshape(-1, ws_dim)
        self.box_half = np.zeros((0, ws_dim), np.float32) if box_half is None else \
            np.ascontiguousarray(box_half, dtype=np.float32).reshape(-1, ws_dim)
        self.cutoff_margin = float(cutoff_margin)
        self.name = name
        assert self.sphere_centers.shape[0] == self.sphere_radii.shape[0]
        assert self.box_centers.shape == self.box_half.shape

    @property
    def n_spheres(self):
        return int(self.sphere_radii.shape[0])

    @property
    def n_boxes(self):
        return int(self.box_centers.shape[0])
