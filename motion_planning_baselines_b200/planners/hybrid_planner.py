"""HybridPlanner: a sample-based planner seeds an optimisation-based planner
(mp_baselines/planners/hybrid_planner.py:10-89).  Same constructor and ``optimize`` signature.  The seed paths are
re-sampled and given velocities in ONE batched device operation (trajectory.smoothen_trajectories) instead of the
reference's per-path loop (:46-62); the optimisation phase runs the fused planner step (``opt_iters=1`` per call,
:77-79).  The sample-based planner itself is duck-typed (``optimize(refill_samples_buffer=..., debug=...) -> list of
[n_i, d] paths or None``, ``start_state_pos``, ``goal_state_pos``): the RRT tree search is out of scope (SURVEY 2,
rows 11-12), its collision query is ``task.PlanningTask.compute_collision``."""
import time

import torch

from ..trajectory import smoothen_trajectories, tensor_linspace
from .base import MPPlanner


class HybridPlanner(MPPlanner):
    def __init__(self, sample_based_planner, opt_based_planner, **kwargs):
        super().__init__('HybridSampleAndOptimizationPlanner', **kwargs)
        self.sample_based_planner = sample_based_planner
        self.opt_based_planner = opt_based_planner

    def render(self, ax, **kwargs):
        raise NotImplementedError

    def optimize(self, debug=False, print_times=False, return_iterations=False, **kwargs):
        opt = self.opt_based_planner
        t0 = time.perf_counter()
        traj_l = self.sample_based_planner.optimize(refill_samples_buffer=True, debug=debug, **kwargs)
        if isinstance(traj_l, torch.Tensor) and traj_l.ndim == 2:
            traj_l = [traj_l]
        t_sample = time.perf_counter() - t0
        if debug or print_times:
            print(f'Sample-based Planner -- Optimization time: {t_sample:.3f} sec')

        # no solution -> straight line between start and goal, even if it is in collision (hybrid_planner.py:48-53)
        paths = []
        for traj in traj_l:
            if traj is None:
                traj = tensor_linspace(self.sample_based_planner.start_state_pos.to(**self.tensor_args),
                                       self.sample_based_planner.goal_state_pos.to(**self.tensor_args),
                                       steps=opt.n_support_points).T
            paths.append(traj.to(**self.tensor_args))
        pos, vel = smoothen_trajectories(paths, n_support_points=opt.n_support_points, dt=opt.dt,
                                         set_average_velocity=True, tensor_args=self.tensor_args)
        initial_traj_pos_vel = torch.cat((pos, vel), dim=-1).unsqueeze(0)        # 'n h d -> 1 n h d': one goal only (:64-66)

        opt.reset(initial_particle_means=initial_traj_pos_vel)
        trajs_0 = opt.get_traj()
        trajs_iters = torch.empty((opt.opt_iters + 1, *trajs_0.shape), **self.tensor_args)
        trajs_iters[0] = trajs_0
        t1 = time.perf_counter()
        for i in range(opt.opt_iters):
            trajs_iters[i + 1] = opt.optimize(opt_iters=1, debug=debug, **kwargs)
        if debug or print_times:
            torch.cuda.synchronize(self.tensor_args['device'])
            print(f'Optimization-based Planner -- Optimization time: {time.perf_counter() - t1:.3f} sec')
            print(f'Hybrid-based Planner -- Optimization time: {time.perf_counter() - t0:.3f} sec')
        return trajs_iters if return_iterations else trajs_iters[-1]
