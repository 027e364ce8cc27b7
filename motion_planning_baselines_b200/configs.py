"""Synthetic workloads for the five BASELINE.json configs.

The reference takes its environments and default hyper-parameters from the un-vendored
``torch_robotics`` package (e.g. examples/pointmass_grid_circles_2d_Stoch-GPMP.py:27-55),
so the obstacle sets below are OURS: seeded, deterministic, published with every result.
Planner parameters that the reference examples do pin are cited next to each value.
"""
import numpy as np

from .models import ObstacleSet, panda_model, point_mass_model


def env_grid_circles_2d(margin=0.005):
    """4x4 grid + offset 3x3 grid of circles in [-1,1]^2 (stand-in for EnvGridCircles2D);
    margin from examples/pointmass_grid_circles_2d_STOMP.py:43."""
    c = []
    for x in np.linspace(-0.75, 0.75, 4):
        for y in np.linspace(-0.75, 0.75, 4):
            if abs(x) > 0.7 and abs(y) > 0.7 and x * y > 0:
                continue                       # keep start (-.8,-.8) / goal (.8,.8) corners free
            c.append((x, y, 0.125))
    for x in np.linspace(-0.5, 0.5, 3):
        for y in np.linspace(-0.5, 0.5, 3):
            c.append((x, y, 0.1))
    c = np.array(c, dtype=np.float32)
    return ObstacleSet(2, sphere_centers=c[:, :2], sphere_radii=c[:, 2], cutoff_margin=margin,
                       name='grid_circles_2d')


def env_dense_2d(margin=0.005, seed=3):
    """16 circles + 8 boxes uniform in [-1,1]^2, sizes U(0.05,0.2)
    (margin: examples/pointmass_dense_2d_CHOMP.py:62)."""
    rng = np.random.default_rng(seed)

    def draw(n):
        out = []
        while len(out) < n:
            p = rng.uniform(-1, 1, 2)
            s = rng.uniform(0.05, 0.2, 2)
            if min(np.linalg.norm(p - np.array([-0.8, -0.8])), np.linalg.norm(p - np.array([0.8, 0.8]))) < 0.35:
                continue
            out.append((p, s))
        return out
    sph, box = draw(16), draw(8)
    return ObstacleSet(2, sphere_centers=[p for p, _ in sph], sphere_radii=[s[0] for _, s in sph],
                       box_centers=[p for p, _ in box], box_half=[s for _, s in box],
                       cutoff_margin=margin, name='dense_2d')


def env_maze_boxes_3d(margin=0.005, seed=0):
    """32 axis-aligned boxes on a jittered 4x4x2 lattice in [-1,1]^3
    (margin: examples/pointmass_maze_boxes_3d_STOMP.py:43)."""
    rng = np.random.default_rng(seed)
    cen, half = [], []
    for x in np.linspace(-0.6, 0.6, 4):
        for y in np.linspace(-0.6, 0.6, 4):
            for z in (-0.4, 0.4):
                cen.append(np.array([x, y, z]) + rng.uniform(-0.08, 0.08, 3))
                half.append(rng.uniform(0.06, 0.14, 3))
    return ObstacleSet(3, box_centers=cen, box_half=half, cutoff_margin=margin, name='maze_boxes_3d')


def env_panda_spheres(margin=0.05, seed=1111):
    """16 spheres, centres U([-1,1]^2 x [0,1]), radii U(0.05,0.15), rejecting the robot base
    column (margin: examples/panda_spheres_GPMP.py:52)."""
    rng = np.random.default_rng(seed)
    c, r = [], []
    while len(c) < 16:
        p = np.array([rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(0, 1)])
        if np.hypot(p[0], p[1]) < 0.35:
            continue
        c.append(p)
        r.append(rng.uniform(0.05, 0.15))
    return ObstacleSet(3, sphere_centers=c, sphere_radii=r, cutoff_margin=margin, name='panda_spheres')


def env_panda_table_shelf(margin=0.03):
    """A table and a 4-board shelf made of 8 boxes (margin: examples/panda_table_shelf_GPMP.py:46)."""
    boxes = [
        ((0.55, 0.0, -0.05), (0.45, 0.6, 0.05)),        # table top
        ((0.0, 0.75, 0.5), (0.3, 0.02, 0.5)),           # shelf back
        ((-0.29, 0.6, 0.5), (0.01, 0.15, 0.5)),         # shelf side
        ((0.29, 0.6, 0.5), (0.01, 0.15, 0.5)),          # shelf side
        ((0.0, 0.6, 0.20), (0.3, 0.15, 0.01)),          # board
        ((0.0, 0.6, 0.45), (0.3, 0.15, 0.01)),          # board
        ((0.0, 0.6, 0.70), (0.3, 0.15, 0.01)),          # board
        ((0.0, 0.6, 0.95), (0.3, 0.15, 0.01)),          # board
    ]
    return ObstacleSet(3, box_centers=[b[0] for b in boxes], box_half=[b[1] for b in boxes],
                       cutoff_margin=margin, name='panda_table_shelf')


# Collision-free joint configurations used as start / goal for the Panda configs
# (inside the limits of models.PANDA_Q_MIN/MAX; cf. examples/panda_spheres_GPMP.py:70-76).
PANDA_START = np.array([0.0, -0.6, 0.0, -2.2, 0.0, 1.6, 0.78], dtype=np.float32)
PANDA_GOAL = np.array([1.2, 0.3, -0.4, -1.4, 0.5, 1.9, -0.3], dtype=np.float32)

STOCH_GPMP_SIGMAS = dict(                      # frozen by us
    sigma_start=1e-3, sigma_gp=1e-1, sigma_goal_prior=1e-3, sigma_coll=1e-4,
    sigma_start_init=1e-3, sigma_goal_init=1e-3, sigma_gp_init=1e-1,
    sigma_start_sample=1e-3, sigma_goal_sample=1e-3, sigma_gp_sample=1e-1,
    temperature=1.0, step_size=0.1,
)


def config(name):
    """-> dict(robot, obstacles, start, goal, H, dt, planner, shape/planner params)."""
    if name == 'C1':   # pointmass_grid_circles_2d_STOMP: 1 x 64 x 64
        return dict(name='pointmass_grid_circles_2d_STOMP', planner='STOMP', robot=point_mass_model(2),
                    obstacles=env_grid_circles_2d(), start=np.array([-0.8, -0.8], np.float32),
                    goal=np.array([0.8, 0.8], np.float32), H=64, dt=0.04, P=1, S=64,
                    params=dict(temperature=1.0, step_size=0.1, sigma_spectral=0.1, sigma_coll=1e-3,
                                sigma_start_init=1e-3, sigma_goal_init=1e-3, sigma_gp_init=5.0))
    if name == 'C2':   # pointmass_dense_2d CHOMP + GPMP: batch 1024 x 64
        return dict(name='pointmass_dense_2d_CHOMP_GPMP', planner='CHOMP+GPMP2', robot=point_mass_model(2),
                    obstacles=env_dense_2d(), start=np.array([-0.8, -0.8], np.float32),
                    goal=np.array([0.8, 0.8], np.float32), H=64, dt=5.0 / 64, P=1024, S=1,
                    params=dict(
                        chomp=dict(weight_prior_cost=1e-4, step_size=0.05, grad_clip=0.05,
                                   sigma_gp_init=0.3, sigma_coll=1.0, cost_weight=10.0, dt=0.04),
                        gpmp2=dict(sigma_start=1e-5, sigma_gp=1e-2, sigma_coll=1e-5, sigma_goal_prior=1e-5,
                                   sigma_start_init=1e-4, sigma_goal_init=1e-4, sigma_gp_init=1e-2,
                                   sigma_start_sample=1e-3, sigma_goal_sample=1e-3,
                                   delta=1e-2, trust_region=True, method='cholesky', step_size=0.5)))
    if name == 'C3':   # pointmass_maze_boxes_3d Stoch-GPMP: 256 x 128 x 64
        return dict(name='pointmass_maze_boxes_3d_StochGPMP', planner='StochGPMP', robot=point_mass_model(3),
                    obstacles=env_maze_boxes_3d(), start=np.array([-0.8, -0.8, -0.8], np.float32),
                    goal=np.array([0.8, 0.8, 0.8], np.float32), H=64, dt=0.04, P=256, S=128,
                    params=dict(STOCH_GPMP_SIGMAS))
    if name == 'C4':   # panda_spheres Stoch-GPMP: 512 x 64 x 64
        return dict(name='panda_spheres_StochGPMP', planner='StochGPMP', robot=panda_model(),
                    obstacles=env_panda_spheres(), start=PANDA_START, goal=PANDA_GOAL,
                    H=64, dt=5.0 / 64, P=512, S=64, params=dict(STOCH_GPMP_SIGMAS))
    if name == 'C5':   # panda_table_shelf MPPI/STOMP sweep
        return dict(name='panda_table_shelf_MPPI', planner='MPPI', robot=panda_model(),
                    obstacles=env_panda_table_shelf(), start=PANDA_START, goal=PANDA_GOAL,
                    H=64, dt=0.04, P=1, S=1000,
                    params=dict(control_std=0.15, temp=1.0, step_size=1.0, cov_prior_type='const_ctrl',
                                c_weights=dict(pos=1., vel=1., ctrl=1., pos_T=1000., vel_T=0.),
                                ctrl_min=-100.0, ctrl_max=100.0, sigma_coll=1e-3))
    raise KeyError(name)
