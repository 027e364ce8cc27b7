"""Generate tests/golden/*.npz by running the UNMODIFIED reference planners.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Run in the build container only:

    python oracle/make_golden.py            # needs /root/reference (read-only) on disk

The reference's ``mp_baselines`` package is imported from /root/reference with the minimal
``torch_robotics`` stand-in under oracle/ref_shim on sys.path; robots and fields are the
duck-typed oracle objects (oracle/robots.py, oracle/fields.py).  All Gaussian noise the
reference draws through torch.distributions.MultivariateNormal is RECORDED (we wrap
``torch.distributions.multivariate_normal._standard_normal``) and stored next to the
results, so the oracle restatement and the CUDA path can be replayed on identical noise.
Nothing here is read at test time on the GPU box: only the committed .npz files are.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get('MPB_REFERENCE', '/root/reference')
sys.path[:0] = [os.path.join(HERE, 'ref_shim'), REF, ROOT]

import torch.distributions.multivariate_normal as _mvn  # noqa: E402

from motion_planning_baselines_b200 import configs, models  # noqa: E402
from oracle.build import oracle_field, oracle_robot  # noqa: E402

TA = dict(device='cpu', dtype=torch.float32)
OUT = os.path.join(ROOT, 'tests', 'golden')


class NoiseRecorder:
    """Context manager recording every standard-normal block MultivariateNormal draws."""

    def __enter__(self):
        self.draws = []
        self._orig = _mvn._standard_normal

        def rec(shape, dtype, device):
            e = self._orig(shape, dtype, device)
            self.draws.append(e.clone())
            return e
        _mvn._standard_normal = rec
        return self

    def __exit__(self, *exc):
        _mvn._standard_normal = self._orig
        return False


def npy(t):
    return t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)


def save(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **{k: npy(v) for k, v in arrays.items()})
    print(f'{name}: {os.path.getsize(path) / 1024:.1f} KiB')


# ---------------------------------------------------------------------------- cases
def gen_stoch_gpmp(tag, cfg_name, P, S, H, iters, sig, seed, store_factor=True):
    from mp_baselines.planners.stoch_gpmp import StochGPMP
    cfg = configs.config(cfg_name)
    model, obst = cfg['robot'], cfg['obstacles']
    robot, field = oracle_robot(model, cfg['dt']), oracle_field(obst, model)
    torch.manual_seed(seed)
    start = torch.tensor(cfg['start'], **TA)
    goal = torch.tensor(cfg['goal'], **TA)
    with NoiseRecorder() as rec:
        planner = StochGPMP(robot=robot, n_dof=model.q_dim, n_support_points=H, num_particles_per_goal=P,
                            opt_iters=1, dt=cfg['dt'], start_state=start, multi_goal_states=goal.unsqueeze(0),
                            collision_fields=[field], tensor_args=TA, num_samples=S, **sig)
        means0 = planner._particle_means.clone()
        L = planner._sample_dist.dist.scale_tril[0].clone()
        n_ctor = len(rec.draws)
        if store_factor:
            out = dict(means0=means0, L=L, Sigma_inv=planner.Sigma_inv, eps_init=rec.draws[0])
        else:
            # H = 64: the [M,M] factor and precision are 3.2 MB each.  The factor couples only entries of the same dof
            # (exact zeros elsewhere, asserted here), so its d [2H,2H] lower-triangular blocks carry all of it; of the
            # precision the main + first 2d sub-diagonals do.  NOTE: at H = 64 the reference's fp32 factorisation is only
            # accurate to ~5e-3 and changes by ~1e-2 with the LAPACK thread count, so a replay needs THIS factor.
            d_, M_ = model.q_dim, L.shape[0]
            blocks = torch.stack([L[j::d_][:, j::d_] for j in range(d_)])
            dense = torch.zeros_like(L)
            for j in range(d_):
                dense[j::d_, j::d_] = blocks[j]
            assert torch.equal(dense, L), 'scale_tril must decouple over the dofs exactly'
            Sinv = planner.Sigma_inv
            out = dict(means0=means0, eps_init=rec.draws[0], L_blocks=blocks,
                       Sinv_band=torch.stack([torch.nn.functional.pad(torch.diagonal(Sinv, -k), (0, k)) for k in range(2 * d_ + 1)]))
        for it in range(iters):
            planner.optimize(opt_iters=1)
            out[f'eps{it}'] = rec.draws[n_ctor + it]
            out[f'samples{it}'] = planner.state_samples
            out[f'weights{it}'] = planner._weights.reshape(P, S)
            out[f'means{it + 1}'] = planner._particle_means.clone()
    # the planner does not keep its costs: re-evaluate them (cost.eval + IS term) on the recorded
    # samples and pre-update means
    for it in range(iters):
        planner.state_samples = torch.as_tensor(out[f'samples{it}'])
        planner._particle_means = torch.as_tensor(out['means0'] if it == 0 else out[f'means{it}']).clone()
        with torch.no_grad():
            out[f'costs{it}'] = planner._get_costs()
            terms, _ = planner.cost.eval(planner.state_samples, return_invidual_costs_and_weights=True)
            out[f'terms{it}'] = torch.stack([t.reshape(-1) for t in terms])
    meta = dict(cfg=cfg_name, P=P, S=S, H=H, d=model.q_dim, dt=cfg['dt'], iters=iters, seed=seed, **sig)
    save(tag, meta=np.array(repr(meta)), start=start, goal=goal, **out)


def gen_prior(tag, d, H, dt, sig_s, sig_gp, sig_g, P, S, seed):
    from mp_baselines.planners.costs.factors.gp_factor import GPFactor
    from mp_baselines.planners.costs.factors.mp_priors_multi import MultiMPPrior
    from mp_baselines.planners.costs.factors.unary_factor import UnaryFactor
    torch.manual_seed(seed)
    start = torch.randn(2 * d, **TA)
    goal = torch.randn(1, 2 * d, **TA)
    K_s = UnaryFactor(2 * d, sig_s, start, TA).K
    K_g = UnaryFactor(2 * d, sig_g, goal[0], TA).K
    Q = GPFactor(d, sig_gp, dt, H - 1, TA).Q_inv[0]
    means = torch.randn(P, H, 2 * d, **TA)
    with NoiseRecorder() as rec:
        prior = MultiMPPrior(H - 1, dt, 2 * d, d, K_s, Q, start, K_g_inv=K_g, means=means,
                             goal_states=goal, tensor_args=TA)
        x = prior.sample(S)
        prior0 = MultiMPPrior(H - 1, dt, 2 * d, d, K_s, Q, start, K_g_inv=K_g, goal_states=goal, tensor_args=TA)
    save(tag, meta=np.array(repr(dict(d=d, H=H, dt=dt, sig_s=sig_s, sig_gp=sig_gp, sig_g=sig_g, P=P, S=S))),
         start=start, goal=goal, K_s=K_s, K_g=K_g, Q_inv=Q, means=means, Sigma_inv=prior.Sigma_inv,
         scale_tril=prior.dist.scale_tril[0], eps=rec.draws[0], samples=x, const_vel_mean=prior0.means)


def gen_stomp(tag, cfg_name, P, S, H, iters, seed, sigma_coll):
    from mp_baselines.planners.costs.cost_functions import CostCollision, CostComposite
    from mp_baselines.planners.stomp import STOMP
    cfg = configs.config(cfg_name)
    model, obst = cfg['robot'], cfg['obstacles']
    robot, field = oracle_robot(model, cfg['dt']), oracle_field(obst, model)
    torch.manual_seed(seed)
    start = torch.tensor(cfg['start'], **TA)
    goal = torch.tensor(cfg['goal'], **TA)
    cost = CostComposite(robot, H, [CostCollision(robot, H, field=field, sigma_coll=sigma_coll, tensor_args=TA)],
                         tensor_args=TA)
    prm = cfg['params'] if cfg_name == 'C1' else configs.config('C1')['params']
    with NoiseRecorder() as rec:
        planner = STOMP(n_dof=model.q_dim, n_support_points=H, num_particles_per_goal=P, num_samples=S,
                        opt_iters=1, dt=cfg['dt'], start_state=start, cost=cost,
                        multi_goal_states=goal.unsqueeze(0), temperature=prm['temperature'],
                        step_size=prm['step_size'], sigma_spectral=prm['sigma_spectral'],
                        sigma_start_init=prm['sigma_start_init'], sigma_goal_init=prm['sigma_goal_init'],
                        sigma_gp_init=prm['sigma_gp_init'], pos_only=False, tensor_args=TA)
        n_ctor = len(rec.draws)
        out = dict(means0=planner._particle_means.clone(), R=planner.Sigma_inv, Sigma=planner.Sigma,
                   L_R=planner._noise_dist.scale_tril.reshape(-1, H, H)[0].clone())
        for it in range(iters):
            planner.optimize(opt_iters=1)
            out[f'eps{it}'] = rec.draws[n_ctor + it]
            out[f'samples{it}'] = planner.state_particles
            out[f'costs{it}'] = planner.costs
            out[f'weights{it}'] = planner._weights.reshape(P, S)
            out[f'means{it + 1}'] = planner._particle_means.clone()
    meta = dict(cfg=cfg_name, P=P, S=S, H=H, d=model.q_dim, dt=cfg['dt'], iters=iters, sigma_coll=sigma_coll,
                temperature=prm['temperature'], step_size=prm['step_size'], sigma_spectral=prm['sigma_spectral'])
    save(tag, meta=np.array(repr(meta)), start=start, goal=goal, **out)


def gen_mppi(tag, N, T, iters, seed, sigma_coll=1e-3):
    from mp_baselines.planners.costs.cost_functions import CostCollision, CostComposite
    from mp_baselines.planners.dynamics.point import PointParticleDynamics
    from mp_baselines.planners.mppi import MPPI
    cfg = configs.config('C1')
    model, obst = cfg['robot'], cfg['obstacles']
    robot, field = oracle_robot(model, cfg['dt']), oracle_field(obst, model)
    torch.manual_seed(seed)
    start = torch.tensor(cfg['start'], **TA)
    goal = torch.tensor(cfg['goal'], **TA)
    cw = dict(pos=1., vel=1., ctrl=1., pos_T=1000., vel_T=0.)
    system = PointParticleDynamics(rollout_steps=T, control_dim=2, state_dim=2, dt=cfg['dt'], discount=1.,
                                   goal_state=goal, ctrl_min=[-100, -100], ctrl_max=[100, 100],
                                   c_weights=cw, tensor_args=TA)
    cost = CostComposite(robot, T, [CostCollision(robot, T, field=field, sigma_coll=sigma_coll, tensor_args=TA)],
                         tensor_args=TA)
    obs = dict(state=start, goal_state=goal, cost=cost)
    with NoiseRecorder() as rec:
        planner = MPPI(system, num_ctrl_samples=N, rollout_steps=T, opt_iters=1, control_std=[0.15, 0.15],
                       temp=1., step_size=1., cov_prior_type='const_ctrl', tensor_args=TA)
        out = dict(mean0=planner._mean.clone(), Cov=planner.ctrl_dist.Cov, Cov_inv=planner.Cov_inv,
                   L_ctrl=torch.stack([dd.scale_tril for dd in planner.ctrl_dist.list_ctrl_dists]))
        for it in range(iters):
            n0 = len(rec.draws)
            U, X, c = planner.optimize(**obs)
            out[f'eps{it}'] = torch.stack(rec.draws[n0:n0 + 2])          # [C,N,T]
            out[f'controls{it}'], out[f'states{it}'], out[f'costs{it}'] = U, X, c
            out[f'weights{it}'] = planner.weights
            out[f'mean{it + 1}'] = planner._mean.clone()
            out[f'best_cost{it}'] = planner.best_cost
            out[f'best_traj{it}'] = planner.best_traj
    meta = dict(N=N, T=T, dt=cfg['dt'], iters=iters, sigma_coll=sigma_coll, c_weights=cw)
    save(tag, meta=np.array(repr(meta)), start=start, goal=goal, **out)


def gen_chomp(tag, cfg_name, P, H, iters, seed, prm):
    from mp_baselines.planners.chomp import CHOMP
    from mp_baselines.planners.costs.cost_functions import CostCollision, CostComposite
    cfg = configs.config(cfg_name)
    model, obst = cfg['robot'], cfg['obstacles']
    robot, field = oracle_robot(model, prm['dt']), oracle_field(obst, model)
    torch.manual_seed(seed)
    start = torch.tensor(cfg['start'], **TA)
    goal = torch.tensor(cfg['goal'], **TA)
    cost = CostComposite(robot, H, [CostCollision(robot, H, field=field, sigma_coll=prm['sigma_coll'], tensor_args=TA)],
                         weights_cost_l=[prm['cost_weight']], tensor_args=TA)
    planner = CHOMP(n_dof=model.q_dim, n_support_points=H, num_particles_per_goal=P, opt_iters=1, dt=prm['dt'],
                    start_state=start, cost=cost, weight_prior_cost=prm['weight_prior_cost'],
                    step_size=prm['step_size'], grad_clip=prm['grad_clip'], multi_goal_states=goal.unsqueeze(0),
                    sigma_start_init=1e-3, sigma_goal_init=1e-3, sigma_gp_init=prm['sigma_gp_init'],
                    pos_only=False, tensor_args=TA)
    out = dict(x0=planner._particle_means.clone(), R=planner.Sigma_inv)
    for it in range(iters):
        planner.optimize(opt_iters=1)
        out[f'x{it + 1}'] = planner._particle_means.clone()
    planner.reset(initial_particle_means=torch.as_tensor(out['x0']))
    planner.optimize(opt_iters=iters)
    out['x_multi'] = planner._particle_means.clone()
    meta = dict(cfg=cfg_name, P=P, H=H, d=model.q_dim, iters=iters, **prm)
    save(tag, meta=np.array(repr(meta)), start=start, goal=goal, **out)


def gen_gpmp2(tag, cfg_name, P, H, iters, seed, prm, dt):
    from mp_baselines.planners.gpmp2 import GPMP2
    cfg = configs.config(cfg_name)
    model, obst = cfg['robot'], cfg['obstacles']
    robot, field = oracle_robot(model, dt), oracle_field(obst, model)
    torch.manual_seed(seed)
    start = torch.tensor(cfg['start'], **TA)
    goal = torch.tensor(cfg['goal'], **TA)
    planner = GPMP2(robot=robot, n_dof=model.q_dim, n_support_points=H, num_particles_per_goal=P, opt_iters=1,
                    dt=dt, start_state=start, multi_goal_states=goal.unsqueeze(0), collision_fields=[field],
                    step_size=prm['step_size'], sigma_start_init=prm['sigma_start_init'],
                    sigma_goal_init=prm['sigma_goal_init'], sigma_gp_init=prm['sigma_gp_init'],
                    sigma_start_sample=prm['sigma_start_sample'], sigma_goal_sample=prm['sigma_goal_sample'],
                    solver_params=dict(delta=prm['delta'], trust_region=prm['trust_region'], method=prm['method']),
                    sigma_start=prm['sigma_start'], sigma_gp=prm['sigma_gp'], sigma_coll=prm['sigma_coll'],
                    sigma_goal_prior=prm['sigma_goal_prior'], tensor_args=TA)
    out = dict(means0=planner._particle_means.clone())
    for it in range(iters):
        A, b, K = planner.cost.get_linear_system(planner._particle_means.clone(), n_interpolated_points=None)
        JtJ, g = planner._get_grad_terms(A, b, K, delta=prm['delta'], trust_region=prm['trust_region'])
        if it == 0:
            out['A0'], out['b0'], out['Kdiag0'] = A[:2], b, torch.diagonal(K, dim1=-2, dim2=-1)
            out['g0'] = g
        planner.optimize(opt_iters=1)
        out[f'means{it + 1}'] = planner._particle_means.clone()
        out[f'costs{it}'] = planner.costs
    meta = dict(cfg=cfg_name, P=P, H=H, d=model.q_dim, dt=dt, iters=iters, **prm)
    save(tag, meta=np.array(repr(meta)), start=start, goal=goal, **out)


MODERATE = dict(sigma_start=1e-2, sigma_gp=1.0, sigma_goal_prior=1e-2, sigma_coll=1e-1,
                sigma_start_init=1e-2, sigma_goal_init=1e-2, sigma_gp_init=1.0,
                sigma_start_sample=1e-2, sigma_goal_sample=1e-2, sigma_gp_sample=1.0,
                temperature=1.0, step_size=0.5)
FROZEN = dict(configs.STOCH_GPMP_SIGMAS)

if __name__ == '__main__':
    torch.set_num_threads(4)
    gen_prior('prior_d2_H16', d=2, H=16, dt=0.04, sig_s=1e-3, sig_gp=1e-1, sig_g=1e-3, P=3, S=5, seed=1)
    gen_prior('prior_d7_H8', d=7, H=8, dt=5 / 64, sig_s=1e-3, sig_gp=1e-1, sig_g=1e-3, P=2, S=4, seed=2)
    gen_stoch_gpmp('stochgpmp_pm2d_moderate', 'C1', P=3, S=8, H=16, iters=2, sig=MODERATE, seed=10)
    gen_stoch_gpmp('stochgpmp_pm3d_moderate', 'C3', P=2, S=8, H=16, iters=2, sig=MODERATE, seed=11)
    gen_stoch_gpmp('stochgpmp_pm3d_frozen', 'C3', P=2, S=8, H=16, iters=1, sig=FROZEN, seed=12)
    gen_stoch_gpmp('stochgpmp_panda_moderate', 'C4', P=2, S=6, H=8, iters=2, sig=MODERATE, seed=13)
    gen_stoch_gpmp('stochgpmp_panda_frozen', 'C4', P=2, S=6, H=8, iters=1, sig=FROZEN, seed=14)
    gen_stomp('stomp_pm2d', 'C1', P=2, S=8, H=16, iters=2, seed=20, sigma_coll=1e-1)
    gen_stomp('stomp_panda', 'C4', P=1, S=6, H=12, iters=1, seed=21, sigma_coll=1e-1)
    gen_mppi('mppi_pm2d', N=16, T=16, iters=2, seed=30)
    c2 = configs.config('C2')['params']
    gen_chomp('chomp_pm2d', 'C2', P=4, H=16, iters=3, seed=40, prm=c2['chomp'])
    gen_gpmp2('gpmp2_pm2d', 'C2', P=3, H=16, iters=2, seed=50, prm=c2['gpmp2'], dt=5 / 64)
    gen_gpmp2('gpmp2_panda', 'C4', P=2, H=8, iters=1, seed=51, prm=c2['gpmp2'], dt=5 / 64)
