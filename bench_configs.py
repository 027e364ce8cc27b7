#!/usr/bin/env python
"""Timings of the other BASELINE.json configs (C1 STOMP, C2 CHOMP + GPMP2, C3 Stoch-GPMP, C5 MPPI / STOMP sweep),
reported as "ms per planner iter" next to bench.py's headline (C4).  Used by bench.py (key "other_configs") and
runnable on its own, also under torchrun for the sample-split variant of C5:

    python bench_configs.py                      # 1 GPU, all configs
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench_configs.py --sample-split          # C5 MPPI / STOMP with ONE problem's samples sharded over N GPUs

Every number goes through the public planner API (planner.optimize), noise drawn on the device by the planner,
timed with CUDA events after 3 warm-up calls."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def _time(fn, iters, warmup=3, sync=None):
    for _ in range(warmup):
        fn()
    if sync is not None:
        sync()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def _timed_split(fn, iters, sync, split, reduce_max):
    """_time + (for a TimedSplit) the stream time of the collectives per iteration, both as max over ranks."""
    for _ in range(3):
        fn()
    if hasattr(split, 'reset'):
        split.reset()
    ms = _time(fn, iters, warmup=0, sync=sync)
    coll = {}
    if hasattr(split, 'collective_ms') and split.world > 1:
        c_ms = split.collective_ms() / iters
        if reduce_max is not None:
            c_ms = reduce_max(c_ms)
        coll = dict(collective_ms_per_iter=c_ms, collectives_per_iter=split.calls // iters,
                    collective_bytes_per_iter=split.bytes // iters)
    if reduce_max is not None:
        ms = reduce_max(ms)
    return dict(ms=ms, collectives=dict(collectives=coll) if coll else {})


def _straight(cfg, P, H, d, dev, jitter=0.0):
    w = torch.linspace(0, 1, H, device=dev['device']).view(1, H, 1)
    x = torch.zeros(P, H, 2 * d, **dev)
    s, g = torch.tensor(cfg['start']).to(**dev), torch.tensor(cfg['goal']).to(**dev)
    x[..., :d] = s * (1 - w) + g * w
    if jitter:
        gen = torch.Generator(device=dev['device']).manual_seed(0)
        x[..., :d] += jitter * torch.randn(P, H, d, generator=gen, **dev).cumsum(1) / H ** 0.5
        x[:, 0, :d], x[:, -1, :d] = s, g
    x[..., d:] = (g - s) / (H * cfg['dt'])
    return x


def _collision_cost(cfg, H, sigma_coll, dev, weight=1.0):
    from motion_planning_baselines_b200.costs import CostCollision, CostComposite
    from motion_planning_baselines_b200.fields import CollisionField
    from motion_planning_baselines_b200.robots import Robot
    robot = Robot(cfg['robot'], dt=cfg['dt'], tensor_args=dev)
    field = CollisionField(cfg['obstacles'], tensor_args=dev)
    return CostComposite(robot, H, [CostCollision(robot, H, field=field, sigma_coll=sigma_coll, tensor_args=dev)],
                         weights_cost_l=[weight], tensor_args=dev)


def launch_floor(dev, n_launches, reps=200):
    """Measured floor of `n_launches` dependent kernel launches on one stream (ms): the smallest kernel of the library
    (mpb_softmax_weights on one element) launched back to back through the same ctypes path, CUDA events around the batch."""
    from motion_planning_baselines_b200 import _lib
    lib = _lib.lib()
    c, l, w = torch.zeros(1, 1, **dev), torch.zeros(1, 2, **dev), torch.zeros(1, 1, **dev)

    def batch():
        for _ in range(n_launches):
            _lib.check(lib.mpb_softmax_weights(_lib.ptr(c), _lib.ptr(l), _lib.ptr(w), 1.0, 1, 1, _lib.stream_ptr()))
    return _time(batch, reps)


def bench_c1_stomp(dev, iters=200):
    from motion_planning_baselines_b200 import configs
    from motion_planning_baselines_b200.planners import STOMP
    cfg = configs.config('C1')
    prm, P, S, H, d = cfg['params'], 1, 64, 64, 2
    planner = STOMP(n_dof=d, n_support_points=H, num_particles_per_goal=P, num_samples=S, opt_iters=1, dt=cfg['dt'],
                    start_state=torch.tensor(cfg['start']).to(**dev), cost=_collision_cost(cfg, H, prm['sigma_coll'], dev),
                    multi_goal_states=torch.tensor(cfg['goal']).to(**dev).unsqueeze(0), temperature=prm['temperature'],
                    step_size=prm['step_size'], sigma_spectral=prm['sigma_spectral'],
                    initial_particle_means=_straight(cfg, P, H, d, dev), pos_only=False, tensor_args=dev)
    ms1 = _time(lambda: planner.optimize(opt_iters=1), iters)
    n_in = 20          # the reference example runs 20 iterations per optimize() call (pointmass_grid_circles_2d_STOMP.py:76)
    ms = _time(lambda: planner.optimize(opt_iters=n_in), max(iters // 4, 5)) / n_in
    floor = launch_floor(dev, 3)
    return dict(config='C1 pointmass_grid_circles_2d STOMP', shape='1 particle x 64 samples x 64 waypoints', ms_per_iter=ms,
                ms_per_single_iter_call=ms1, samples_per_s=P * S / (ms * 1e-3), launch_floor_ms_per_iter=floor,
                bound='launch latency', note=f'{n_in} iterations per optimize() call, enqueued by one mpb_stomp_run call (3 launches '
                'per iteration, no host work in between); launch_floor_ms_per_iter = 3 back-to-back launches of the smallest kernel '
                'of the library (mpb_softmax_weights on one element) measured the same way')


def bench_c2(dev, iters=20):
    from motion_planning_baselines_b200 import configs
    from motion_planning_baselines_b200.fields import CollisionField
    from motion_planning_baselines_b200.planners import CHOMP, GPMP2
    from motion_planning_baselines_b200.robots import Robot
    cfg = configs.config('C2')
    P, H, d = 1024, 64, 2
    pc, pg = cfg['params']['chomp'], cfg['params']['gpmp2']
    x0 = _straight(cfg, P, H, d, dev, jitter=0.3)
    cfg_c = dict(cfg, dt=pc['dt'])
    chomp = CHOMP(n_dof=d, n_support_points=H, num_particles_per_goal=P, opt_iters=1, dt=pc['dt'],
                  start_state=torch.tensor(cfg['start']).to(**dev), cost=_collision_cost(cfg_c, H, pc['sigma_coll'], dev, pc['cost_weight']),
                  weight_prior_cost=pc['weight_prior_cost'], step_size=pc['step_size'], grad_clip=pc['grad_clip'],
                  multi_goal_states=torch.tensor(cfg['goal']).to(**dev).unsqueeze(0), initial_particle_means=x0,
                  pos_only=False, tensor_args=dev)
    n_in = 50
    ms_c = _time(lambda: chomp.optimize(opt_iters=n_in), iters) / n_in
    ms_c1 = _time(lambda: chomp.optimize(opt_iters=1), iters)
    robot = Robot(cfg['robot'], dt=cfg['dt'], tensor_args=dev)
    gp = GPMP2(robot=robot, n_dof=d, n_support_points=H, num_particles_per_goal=P, opt_iters=1, dt=cfg['dt'],
               start_state=torch.tensor(cfg['start']).to(**dev), multi_goal_states=torch.tensor(cfg['goal']).to(**dev).unsqueeze(0),
               collision_fields=[CollisionField(cfg['obstacles'], tensor_args=dev)], step_size=pg['step_size'],
               sigma_start_init=pg['sigma_start_init'], sigma_goal_init=pg['sigma_goal_init'], sigma_gp_init=pg['sigma_gp_init'],
               sigma_start_sample=pg['sigma_start_sample'], sigma_goal_sample=pg['sigma_goal_sample'],
               solver_params=dict(delta=pg['delta'], trust_region=pg['trust_region'], method=pg['method']),
               sigma_start=pg['sigma_start'], sigma_gp=pg['sigma_gp'], sigma_coll=pg['sigma_coll'],
               sigma_goal_prior=pg['sigma_goal_prior'], initial_particle_means=x0.unsqueeze(0), tensor_args=dev)
    ms_g = _time(lambda: gp.optimize(opt_iters=1), iters)
    return [dict(config='C2 pointmass_dense_2d CHOMP', shape='1024 trajectories x 64 waypoints', ms_per_iter=ms_c,
                 ms_per_single_iter_call=ms_c1, trajectories_per_s=P / (ms_c * 1e-3),
                 note=f'{n_in} iterations inside one mpb_chomp_run launch; ms_per_single_iter_call = optimize(opt_iters=1)'),
            dict(config='C2 pointmass_dense_2d GPMP2', shape='1024 trajectories x 64 waypoints', ms_per_iter=ms_g,
                 trajectories_per_s=P / (ms_g * 1e-3), launch_floor_ms_per_iter=launch_floor(dev, 3),
                 note='linearize + batch-mean diagonal + block-tridiagonal fp64 Cholesky solve (3 launches)')]


def bench_stoch(dev, name, iters=30):
    from motion_planning_baselines_b200 import configs
    from motion_planning_baselines_b200.fields import CollisionField
    from motion_planning_baselines_b200.planners import StochGPMP
    from motion_planning_baselines_b200.robots import Robot
    cfg = configs.config(name)
    d = cfg['robot'].q_dim
    P, S, H = cfg['P'], cfg['S'], cfg['H']
    robot = Robot(cfg['robot'], dt=cfg['dt'], tensor_args=dev)
    planner = StochGPMP(robot=robot, n_dof=d, n_support_points=H, num_particles_per_goal=P, opt_iters=1, dt=cfg['dt'],
                        start_state=torch.tensor(cfg['start']).to(**dev), multi_goal_states=torch.tensor(cfg['goal']).to(**dev).unsqueeze(0),
                        collision_fields=[CollisionField(cfg['obstacles'], tensor_args=dev)], tensor_args=dev, num_samples=S,
                        **cfg['params'])
    ms = _time(lambda: planner.optimize(opt_iters=1), iters)
    return dict(config=f'{name} {cfg["name"]}', shape=f'{P} particles x {S} samples x {H} waypoints', ms_per_iter=ms,
                samples_per_s=P * S / (ms * 1e-3))


class TimedSplit:
    """SampleSplit whose collectives are bracketed by CUDA events on the launching stream, so that a sample-split run
    reports how long its exchange step occupies the stream per iteration (waiting for the slowest peer included)."""

    def __new__(cls, *a, **kw):
        from motion_planning_baselines_b200.update import SampleSplit

        class _Timed(SampleSplit):
            def __init__(self, *a, **kw):
                super().__init__(*a, **kw)
                self.events, self.calls, self.bytes = [], 0, 0

            def all_gather_cat(self, t):
                if self.world == 1:
                    return t
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                out = super().all_gather_cat(t)
                e1.record()
                self.events.append((e0, e1))
                self.calls += 1
                self.bytes += out.numel() * out.element_size()
                return out

            def reset(self):
                self.events, self.calls, self.bytes = [], 0, 0

            def collective_ms(self):
                torch.cuda.synchronize()
                return sum(a.elapsed_time(b) for a, b in self.events)
        return _Timed(*a, **kw)


def bench_c5(dev, Ns=(1000, 10000, 100000, 1000000), split=None, iters=10, stomp_Ns=None, reduce_max=None):
    """MPPI at every N of ``Ns`` and STOMP at ``stomp_Ns`` (default: the first three).  With a (Timed)SampleSplit the N
    samples of the ONE problem are sharded over the ranks; ``reduce_max`` turns a rank-local time into the max over ranks."""
    from motion_planning_baselines_b200 import configs
    from motion_planning_baselines_b200.costs import CostCollision, CostComposite
    from motion_planning_baselines_b200.dynamics import PointParticleDynamics
    from motion_planning_baselines_b200.fields import CollisionField
    from motion_planning_baselines_b200.planners import MPPI, STOMP
    from motion_planning_baselines_b200.robots import Robot
    cfg = configs.config('C5')
    prm, Tn, d = cfg['params'], 64, 7
    world = split.world if split is not None else 1
    sync = (lambda: torch.distributed.barrier()) if world > 1 else None
    out = []
    robot = Robot(cfg['robot'], dt=cfg['dt'], tensor_args=dev)
    field = CollisionField(cfg['obstacles'], tensor_args=dev)
    cost = CostComposite(robot, Tn, [CostCollision(robot, Tn, field=field, sigma_coll=prm['sigma_coll'], tensor_args=dev)], tensor_args=dev)
    obs = dict(state=torch.tensor(cfg['start']).to(**dev), goal_state=torch.tensor(cfg['goal']).to(**dev), cost=cost)
    for N in Ns:
        system = PointParticleDynamics(rollout_steps=Tn, control_dim=d, state_dim=d, dt=cfg['dt'], discount=1.,
                                       goal_state=torch.tensor(cfg['goal']), ctrl_min=[prm['ctrl_min']] * d,
                                       ctrl_max=[prm['ctrl_max']] * d, c_weights=prm['c_weights'], tensor_args=dev)
        planner = MPPI(system, num_ctrl_samples=N, rollout_steps=Tn, opt_iters=1, control_std=prm['control_std'], temp=prm['temp'],
                       step_size=prm['step_size'], cov_prior_type=prm['cov_prior_type'], tensor_args=dev, sample_split=split)
        n_it = iters if N <= 100000 else 5
        ms = _timed_split(lambda: planner.optimize(opt_iters=1, **obs), n_it, sync, split, reduce_max)
        out.append(dict(config='C5 panda_table_shelf MPPI', shape=f'{N} control samples x 64 steps x 7 dof', n_gpus=world,
                        ms_per_iter=ms['ms'], samples_per_s=N / (ms['ms'] * 1e-3),
                        sharding='samples split over ranks; all-gather of packed records + fixed-order combine' if world > 1 else 'single GPU',
                        **ms['collectives']))
        del planner
        torch.cuda.empty_cache()
    c1 = configs.config('C1')['params']
    for S in (Ns[:3] if stomp_Ns is None else stomp_Ns):
        planner = STOMP(n_dof=d, n_support_points=Tn, num_particles_per_goal=1, num_samples=S, opt_iters=1, dt=cfg['dt'],
                        start_state=torch.tensor(cfg['start']).to(**dev), cost=cost,
                        multi_goal_states=torch.tensor(cfg['goal']).to(**dev).unsqueeze(0), temperature=c1['temperature'],
                        step_size=c1['step_size'], sigma_spectral=c1['sigma_spectral'],
                        initial_particle_means=_straight(cfg, 1, Tn, d, dev), pos_only=False, tensor_args=dev, sample_split=split)
        ms = _timed_split(lambda: planner.optimize(opt_iters=1), iters, sync, split, reduce_max)
        out.append(dict(config='C5 panda_table_shelf STOMP', shape=f'1 particle x {S} samples x 64 waypoints', n_gpus=world,
                        ms_per_iter=ms['ms'], samples_per_s=S / (ms['ms'] * 1e-3), **ms['collectives']))
        del planner
        torch.cuda.empty_cache()
    return out


def run_other_configs(dev, quick=False):
    """-> list of per-config dicts (single GPU)."""
    res = [bench_c1_stomp(dev, iters=50 if quick else 200)]
    res += bench_c2(dev, iters=5 if quick else 20)
    res.append(bench_stoch(dev, 'C3', iters=10 if quick else 30))
    res += bench_c5(dev, Ns=(1000, 10000, 100000, 1000000), stomp_Ns=(1000, 10000, 100000), iters=5 if quick else 10)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--sample-split', action='store_true')
    ap.add_argument('--quick', action='store_true')
    args = ap.parse_args()
    world, rank, local = int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('RANK', '0')), int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = dict(device=torch.device('cuda', local), dtype=torch.float32)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    if args.sample_split:
        res = bench_c5(dev, split=TimedSplit())
    else:
        res = run_other_configs(dev, quick=args.quick)
    if rank == 0:
        for r in res:
            print(json.dumps(r), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
