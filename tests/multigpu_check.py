"""Multi-GPU check of the sample-split mode, launched by torchrun (one process per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py

Every rank also runs the SAME problem unsplit on its own GPU with the same injected noise; the split run must give
the same costs (bit for bit: the per-sample work is identical), the same global argmin, and means that agree to 1e-5
and are bit-identical ACROSS ranks (fixed-order combine of the all-gathered records)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from motion_planning_baselines_b200 import configs  # noqa: E402
from motion_planning_baselines_b200.update import SampleSplit  # noqa: E402


def close(a, b, rtol=1e-5, atol=1e-6):
    return bool(((a.double() - b.double()).abs() <= atol + rtol * b.double().abs()).all())


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    dev = dict(device=torch.device('cuda', local), dtype=torch.float32)
    from test_gpu_mppi import _mppi
    from test_gpu_planners import _collision_cost
    from motion_planning_baselines_b200.planners import STOMP
    split = SampleSplit()
    assert (split.rank, split.world) == (rank, world)
    ok = True

    # ---- MPPI, Panda table+shelf, N = 10 001 control samples (ragged split) -------------------------------------
    cfg = configs.config('C5')
    N, Tn, d = 10001, 64, 7
    cw = cfg['params']['c_weights']
    gen = torch.Generator(device='cuda').manual_seed(123)            # same noise on every rank
    p_split, cost = _mppi(cfg, N, Tn, dev, 1e-1, cw, cfg['params']['control_std'], split=split)
    p_full, cost_f = _mppi(cfg, N, Tn, dev, 1e-1, cw, cfg['params']['control_std'])
    obs = dict(state=torch.tensor(cfg['start']).to(**dev), goal_state=torch.tensor(cfg['goal']).to(**dev))
    for it in range(2):
        eps = torch.randn(d, N, Tn, generator=gen, **dev)
        U, X, c = p_split.optimize(opt_iters=1, eps=[eps], cost=cost, **obs)
        Uf, Xf, cf = p_full.optimize(opt_iters=1, eps=[eps], cost=cost_f, **obs)
        off, cnt = split.local_slice(N)
        gathered = split.all_gather_cat(p_split._mean.unsqueeze(0).contiguous())
        checks = dict(
            rollouts=torch.equal(X, Xf[off:off + cnt]) and torch.equal(U, Uf[off:off + cnt]),
            costs=close(c, cf[off:off + cnt], rtol=1e-6, atol=0),          # the batch sum is reduced in a different order
            argmin=int(p_split._best[1][0]) == int(p_full._best[1][0]),
            mean=close(p_split._mean, p_full._mean),
            best_traj=torch.equal(p_split.best_traj, p_full.best_traj),
            best_cost=close(p_split.best_cost, p_full.best_cost, rtol=1e-6, atol=0),
            identical_across_ranks=all(torch.equal(gathered[0], gathered[r]) for r in range(world)))
        if not all(checks.values()):
            print(f'[rank {rank}] MPPI iter {it}: {checks} energy split {float(p_split._energy):.9g} full {float(p_full._energy):.9g} '
                  f'max|dmean| {float((p_split._mean - p_full._mean).abs().max()):.3e}', flush=True)
        ok &= all(checks.values())
        p_full._mean.copy_(p_split._mean)
        p_full.update_ctrl_dist()      # controls are sampled around ctrl_dist.mu (as in the reference)
    print(f'[rank {rank}] MPPI sample-split x{world}: {"ok" if ok else "MISMATCH"}', flush=True)

    # ---- STOMP, 2-D point mass, one particle with 4 099 samples ------------------------------------------------------
    cfg = configs.config('C1')
    P, S, H, dd = 1, 4099, 64, 2
    prm = cfg['params']
    w = torch.linspace(0, 1, H).view(1, H, 1)
    means = torch.zeros(P, H, 2 * dd)
    means[..., :dd] = torch.tensor(cfg['start']) * (1 - w) + torch.tensor(cfg['goal']) * w
    mk = lambda sp: STOMP(n_dof=dd, n_support_points=H, num_particles_per_goal=P, num_samples=S, opt_iters=1, dt=cfg['dt'],
                          start_state=torch.tensor(cfg['start']).to(**dev), cost=_collision_cost(cfg, H, 1e-1, dev),
                          multi_goal_states=torch.tensor(cfg['goal']).to(**dev).unsqueeze(0), temperature=prm['temperature'],
                          step_size=prm['step_size'], sigma_spectral=prm['sigma_spectral'],
                          initial_particle_means=means.to(**dev), pos_only=False, tensor_args=dev, sample_split=sp)
    s_split, s_full = mk(split), mk(None)
    ok2 = True
    for it in range(2):
        eps = torch.randn(S, 2 * dd, P, H, generator=gen, **dev)
        t1 = s_split.optimize(opt_iters=1, eps=[eps])
        t2 = s_full.optimize(opt_iters=1, eps=[eps])
        off, cnt = split.local_slice(S)
        ok2 &= torch.equal(s_split.costs, s_full.costs[:, off:off + cnt])
        ok2 &= close(t1, t2)
        gathered = split.all_gather_cat(t1.unsqueeze(0).contiguous())
        ok2 &= all(torch.equal(gathered[0], gathered[r]) for r in range(world))
        s_full._particle_means.copy_(s_split._particle_means)
    print(f'[rank {rank}] STOMP sample-split x{world}: {"ok" if ok2 else "MISMATCH"}', flush=True)
    # ---- GPMP2 (trust region = batch mean, quirk B10) and CHOMP (global-P smoothness, quirk B1): particles sharded ------
    from test_gpu_planners import _chomp, _gpmp2
    from conftest import load_golden
    import numpy as np
    g = load_golden('gpmp2_pm2d')
    Bg, H, dd = 64 * world, 32, 2
    cfg = configs.config('C2')
    gen_c = torch.Generator().manual_seed(5)
    w = torch.linspace(0, 1, H).view(1, H, 1)
    xg = torch.zeros(Bg, H, 2 * dd)
    xg[..., :dd] = torch.tensor(cfg['start']) * (1 - w) + torch.tensor(cfg['goal']) * w + 0.25 * torch.randn(Bg, H, dd, generator=gen_c).cumsum(1) / np.sqrt(H)
    off, cnt = split.local_slice(Bg)
    full = _gpmp2(g, dev, means=xg, P=Bg, H=H)
    part = _gpmp2(g, dev, means=xg[off:off + cnt], P=cnt, H=H)
    part.batch_split = split
    ok3 = True
    for it in range(2):
        tf = full.optimize(opt_iters=1)
        tp = part.optimize(opt_iters=1)
        ok3 &= close(tp, tf[off:off + cnt], rtol=1e-6, atol=1e-7)
    gc = load_golden('chomp_pm2d')
    gc2 = dict(gc, meta=dict(gc['meta'], H=H))
    cf = _chomp(gc2, dev, P=Bg, x0=xg)
    cp = _chomp(gc2, dev, P=cnt, x0=xg[off:off + cnt])
    cp.num_particles_global = Bg
    ok3 &= torch.equal(cp.optimize(opt_iters=5), cf.optimize(opt_iters=5)[off:off + cnt])
    print(f'[rank {rank}] GPMP2 / CHOMP particle-split x{world}: {"ok" if ok3 else "MISMATCH"}', flush=True)
    flag = torch.tensor([int(ok and ok2 and ok3)], device=dev['device'])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print('MULTIGPU_CHECK ' + ('PASS' if int(flag) else 'FAIL'), flush=True)
    sys.exit(0 if int(flag) else 1)


if __name__ == '__main__':
    main()
