"""In-kernel noise (csrc/philox.cuh): Philox4x32-10 keyed on the GLOBAL element index.

* the device generator against the numpy restatement (oracle/philox.py, pinned by the Random123 known-answer vectors);
* layouts: a sharded call draws exactly its slice of the job's global stream (world 1 / 2 / 4 splits identical);
* the fused samplers (K1 structured, STOMP, MPPI) with noise drawn in the kernel are BIT-IDENTICAL to the same kernels fed
  with mpb_philox_normal's dump -- so a run can be replayed through the oracle on identical noise;
* moments / Kolmogorov-Smirnov of the stream;
* planner-level: optimize() without injected noise (the way the reference is called, stoch_gpmp.py:281-309) replayed
  through the CPU oracle at 1e-5; same seed => same result, next draw => different noise;
* MPPI pop()/shift(): controls sampled around the UNSHIFTED mean as in the reference (mppi.py:68-70,171-178).
"""
import ctypes as C

import numpy as np
import pytest
import torch

from test_gpu_stoch_gpmp import T, assert_close

pytestmark = pytest.mark.gpu

from motion_planning_baselines_b200 import _lib, configs  # noqa: E402


@pytest.fixture(scope='module')
def dev():
    return dict(device=torch.device('cuda:0'), dtype=torch.float32)


def nd(seed=1, offset=0, s_off=0, p_off=0, P_glob=1):
    return _lib.NoiseDesc(seed=seed, offset=offset, s_offset=s_off, p_offset=p_off, P_global=P_glob)


def assert_normals_close(got, ref):
    """Device (MUFU lg2 / sin / cos) vs the float64 evaluation of the same fp32 inputs (oracle/philox.py): the radius
    r = sqrt(-2 ln a) carries lg2's absolute error 2^-22 amplified by 1 / r, the angle sin/cos's 2^-21.4 times r."""
    n = ref.size - ref.size % 2
    r = np.sqrt(ref[0:n:2] ** 2 + ref[1:n:2] ** 2).repeat(2)       # pairs (cos, sin) share their radius
    err = np.abs(got[:n] - ref[:n])
    bound = 2e-6 + 1e-6 * r + 4e-7 / np.maximum(r, 1e-4)
    assert (err <= bound).all(), f'max err {err.max():.3e} at r = {r[err.argmax()]:.3e}'


def test_generator_vs_numpy_oracle(dev):
    from oracle.philox import philox_normal
    # SPM layout with P_glob = P, no offsets: local [S,P,M] flat index == global element index
    for seed, offset, (S, P, M) in [(1, 0, (3, 2, 8)), (0xDEADBEEFCAFEF00D, (1 << 40) + 5, (5, 3, 28)), (77, 2, (2, 2, 7))]:
        got = _lib.philox_normal(nd(seed, offset, P_glob=P), _lib.NOISE_SPM, (S, P, M), dev['device']).cpu().double().numpy().ravel()
        ref = philox_normal(seed, offset, 0, S * P * M)
        assert_normals_close(got, ref)
    # far into the stream (group index beyond 2^32)
    big = nd(5, 9, s_off=(1 << 36), P_glob=3)
    got = _lib.philox_normal(big, _lib.NOISE_SPM, (2, 3, 8), dev['device']).cpu().double().numpy().ravel()
    ref = philox_normal(5, 9, (1 << 36) * 3 * 8, 2 * 3 * 8)
    assert_normals_close(got, ref)
    # a large block: the bound holds over 2M draws (incl. the tails and the near-zero radii)
    got = _lib.philox_normal(nd(9, 1, P_glob=64), _lib.NOISE_SPM, (128, 64, 256), dev['device']).cpu().double().numpy().ravel()
    assert_normals_close(got, philox_normal(9, 1, 0, got.size))


@pytest.mark.parametrize('layout,shape', [(0, (6, 8, 12)), (1, (6, 3, 8, 16)), (2, (3, 8, 12)), (1, (4, 2, 4, 7))])
def test_shards_draw_their_slice_of_the_global_stream(layout, shape, dev):
    """world 1 vs 2 vs 4: particle sharding (SPM, STOMP) and sample splitting (all layouts)."""
    d = dev['device']
    if layout == _lib.NOISE_SPM:
        S, P, M = shape
        whole = _lib.philox_normal(nd(3, 1, P_glob=P), layout, shape, d)
        for w in (2, 4):
            parts = [_lib.philox_normal(nd(3, 1, p_off=r * P // w, P_glob=P), layout, (S, P // w, M), d) for r in range(w)]
            assert torch.equal(torch.cat(parts, dim=1), whole)
            parts = [_lib.philox_normal(nd(3, 1, s_off=r * S // 2, P_glob=P), layout, (S // 2, P, M), d) for r in range(2)]
            assert torch.equal(torch.cat(parts, dim=0), whole)
    elif layout == _lib.NOISE_STOMP:
        S, D, P, H = shape
        whole = _lib.philox_normal(nd(3, 1, P_glob=P), layout, shape, d)
        for w in (2, 4):
            parts = [_lib.philox_normal(nd(3, 1, p_off=r * P // w, P_glob=P), layout, (S, D, P // w, H), d) for r in range(w)]
            assert torch.equal(torch.cat(parts, dim=2), whole)
        parts = [_lib.philox_normal(nd(3, 1, s_off=r * S // 2, P_glob=P), layout, (S // 2, D, P, H), d) for r in range(2)]
        assert torch.equal(torch.cat(parts, dim=0), whole)
    else:
        Cc, N, Tn = shape
        whole = _lib.philox_normal(nd(3, 1, P_glob=N), layout, shape, d)
        for w in (2, 4):
            parts = [_lib.philox_normal(nd(3, 1, s_off=r * N // w, P_glob=N), layout, (Cc, N // w, Tn), d) for r in range(w)]
            assert torch.equal(torch.cat(parts, dim=1), whole)


def test_moments_and_ks(dev):
    n = _lib.philox_normal(nd(2024, 0, P_glob=64), _lib.NOISE_SPM, (256, 64, 256), dev['device']).double().flatten()
    N = n.numel()
    assert abs(float(n.mean())) < 4 / np.sqrt(N) and abs(float(n.var()) - 1) < 6 * np.sqrt(2 / N)
    assert abs(float((n ** 3).mean())) < 6 * np.sqrt(15 / N) and abs(float((n ** 4).mean()) - 3) < 6 * np.sqrt(96 / N)
    sub = n[:200000].sort().values.cpu()
    cdf = 0.5 * (1 + torch.erf(sub / np.sqrt(2)))
    k = torch.arange(1, sub.numel() + 1, dtype=torch.float64) / sub.numel()
    ks = float(torch.max((cdf - k).abs().max(), (cdf - k + 1 / sub.numel()).abs().max()))
    assert ks < 1.95 / np.sqrt(sub.numel()), f'KS statistic {ks:.2e}'        # 99.9 % level
    # neighbouring elements / rows are uncorrelated
    a = n.view(256 * 64, 256)
    assert abs(float((a[:, :-1] * a[:, 1:]).mean())) < 5 / np.sqrt(N)
    assert abs(float((a[:-1] * a[1:]).mean())) < 5 / np.sqrt(N)


@pytest.mark.parametrize('P,S', [(8, 64), (5, 24), (3, 7)])
def test_k1_fused_noise_is_bit_identical_to_dump_then_sample(P, S, dev, monkeypatch):
    from motion_planning_baselines_b200.factors import GPFactor, MultiMPPrior, UnaryFactor
    monkeypatch.setenv('MPB_SAMPLE_GP', 'kron')      # the warp-MMA sampler draws and reads noise in ONE kernel body: bit identity

    d, H, dt = 7, 64, 5 / 64
    K = UnaryFactor(2 * d, 1e-3, None, dev).K
    Q = GPFactor(d, 1e-1, dt, H - 1, dev).Q_inv[0]
    gen = torch.Generator().manual_seed(0)
    means = torch.randn(P, H, 2 * d, generator=gen).to(**dev)
    prior = MultiMPPrior(H - 1, dt, 2 * d, d, K, Q, torch.zeros(2 * d, **dev), means=means, K_g_inv=K,
                         goal_states=torch.zeros(1, 2 * d, **dev), tensor_args=dev)
    assert prior.kron_tc_kind == 1
    desc = nd(11, 4, p_off=3, P_glob=P + 5)
    x_rng = prior.sample(S, noise_desc=desc).clone()
    eps = _lib.philox_normal(desc, _lib.NOISE_SPM, (S, P, H * 2 * d), dev['device'])
    x_eps = prior.sample(S, eps=eps)
    assert torch.equal(x_rng, x_eps), 'in-kernel noise must reproduce the dumped noise bit for bit'
    # sharding the particles over 2 "ranks" gives the same samples
    if P % 2 == 0:
        halves = []
        for r in range(2):
            pr = MultiMPPrior(H - 1, dt, 2 * d, d, K, Q, torch.zeros(2 * d, **dev), means=means[r * P // 2:(r + 1) * P // 2],
                              K_g_inv=K, goal_states=torch.zeros(1, 2 * d, **dev), tensor_args=dev)
            halves.append(pr.sample(S, noise_desc=nd(11, 4, p_off=3 + r * P // 2, P_glob=P + 5)).clone())
        assert torch.equal(torch.cat(halves, dim=0), x_rng)
    # the default stream advances: two draws differ, same seed repeats
    a = prior.sample(S).clone()
    b = prior.sample(S).clone()
    assert not torch.equal(a, b)
    prior.noise.offset = 0
    assert torch.equal(prior.sample(S), a)


def test_stomp_and_mppi_fused_noise_bit_identical(dev):
    lib = _lib.lib()
    P, S, H, D = 3, 10, 16, 4
    gen = torch.Generator().manual_seed(1)
    LR = torch.tril(torch.randn(H, H, generator=gen)).to(**dev).contiguous()
    mu = torch.randn(P, H, D, generator=gen).to(**dev)
    for (s_off, p_off, Pg) in [(0, 0, P), (7, 2, P + 4)]:
        desc = nd(5, 2, s_off=s_off, p_off=p_off, P_glob=Pg)
        x1, x2 = torch.empty(P, S, H, D, **dev), torch.empty(P, S, H, D, **dev)
        _lib.check(lib.mpb_sample_stomp_rng(_lib.ptr(LR), _lib.ptr(mu), C.byref(desc), _lib.ptr(x1), P, S, H, D, _lib.stream_ptr()))
        eps = _lib.philox_normal(desc, _lib.NOISE_STOMP, (S, D, P, H), dev['device'])
        _lib.check(lib.mpb_sample_stomp(_lib.ptr(LR), _lib.ptr(mu), _lib.ptr(eps), _lib.ptr(x2), P, S, H, D, _lib.stream_ptr()))
        assert torch.equal(x1, x2)
    # H not a multiple of 4: element-wise path
    H2 = 7
    LR2 = torch.tril(torch.randn(H2, H2, generator=gen)).to(**dev).contiguous()
    mu2 = torch.randn(P, H2, D, generator=gen).to(**dev)
    desc = nd(5, 3, P_glob=P)
    x1, x2 = torch.empty(P, S, H2, D, **dev), torch.empty(P, S, H2, D, **dev)
    _lib.check(lib.mpb_sample_stomp_rng(_lib.ptr(LR2), _lib.ptr(mu2), C.byref(desc), _lib.ptr(x1), P, S, H2, D, _lib.stream_ptr()))
    eps = _lib.philox_normal(desc, _lib.NOISE_STOMP, (S, D, P, H2), dev['device'])
    _lib.check(lib.mpb_sample_stomp(_lib.ptr(LR2), _lib.ptr(mu2), _lib.ptr(eps), _lib.ptr(x2), P, S, H2, D, _lib.stream_ptr()))
    assert torch.equal(x1, x2)


def test_mppi_planner_noise_and_shift_quirk(dev):
    """MPPI.optimize() without injected noise == the same call on the dumped noise; after pop() the controls are sampled
    around the UNSHIFTED mean while the IS term / update use the shifted one (reference mppi.py:68-70,125-128,171-178)."""
    from oracle import planners as op
    from test_gpu_mppi import _mppi
    cfg = configs.config('C1')
    N, Tn, d = 96, 16, 2
    cw = dict(pos=1., vel=1., ctrl=1., pos_T=1000., vel_T=0.)
    torch.manual_seed(123)
    planner, cost = _mppi(cfg, N, Tn, dev, 1e-1, cw, [0.15, 0.15])
    planner2, _ = _mppi(cfg, N, Tn, dev, 1e-1, cw, [0.15, 0.15])
    obs = dict(state=T(cfg['start']).to(**dev), goal_state=T(cfg['goal']).to(**dev))
    start, goal = T(cfg['start']), T(cfg['goal'])
    lo, hi = torch.full((d,), -100.), torch.full((d,), 100.)
    assert planner._noise.seed == 123
    for it in range(3):
        desc = planner._noise.desc()
        mean_is, mean_s = planner._mean.clone().cpu(), planner.ctrl_dist.mu.clone().cpu()
        U, X, c = planner.optimize(opt_iters=1, **obs)
        eps = _lib.philox_normal(desc, _lib.NOISE_MPPI, (d, N, Tn), dev['device'])
        planner2._mean = mean_is.to(**dev).clone()
        planner2.ctrl_dist.mu = mean_s.to(**dev).clone()
        U2, X2, c2 = planner2.optimize(opt_iters=1, eps=[eps], **obs)
        assert torch.equal(U, U2) and torch.equal(X, X2) and torch.equal(c, c2)
        ref = op.mppi_iteration(mean_is, planner.ctrl_dist.scale_tril.cpu(), planner.Cov_inv.cpu(), eps.cpu(), start, goal,
                                cfg['dt'], lo, hi, cw, 1.0, 1.0, mean_sample=mean_s)
        assert_close(U, ref['controls'], rtol=1e-5, atol=1e-6, what=f'controls, iteration {it}')
        assert_close(c, ref['costs'], rtol=2e-5, what=f'costs, iteration {it}')
        assert int(c.argmin()) == int(ref['argmin'])
        if it == 0:
            assert torch.equal(mean_is, mean_s)
        else:
            assert not torch.equal(mean_is, mean_s), 'after pop() the sampling mean must lag the shifted mean'
        planner.pop()          # receding horizon: shift() does NOT refresh ctrl_dist (reference behaviour)


@pytest.mark.parametrize('cfg_name,P,S,H', [('C4', 6, 16, 64), ('C3', 4, 32, 32)])
def test_planner_default_noise_replays_through_the_oracle(cfg_name, P, S, H, dev):
    from motion_planning_baselines_b200.fields import CollisionField
    from motion_planning_baselines_b200.planners import StochGPMP
    from motion_planning_baselines_b200.robots import Robot
    from oracle import check
    cfg = configs.config(cfg_name)
    sig = dict(sigma_start=1e-2, sigma_gp=1.0, sigma_goal_prior=1e-2, sigma_coll=1e-1, sigma_start_init=1e-2,
               sigma_goal_init=1e-2, sigma_gp_init=1.0, sigma_start_sample=1e-2, sigma_goal_sample=1e-2,
               sigma_gp_sample=1.0, temperature=1.0, step_size=0.5)
    d = cfg['robot'].q_dim

    def make(seed, p_off=0, P_loc=P, P_glob=P):
        robot = Robot(cfg['robot'], dt=cfg['dt'], tensor_args=dev)
        return StochGPMP(robot=robot, n_dof=d, n_support_points=H, num_particles_per_goal=P_loc, opt_iters=1, dt=cfg['dt'],
                         start_state=T(cfg['start']).to(**dev), multi_goal_states=T(cfg['goal']).to(**dev).unsqueeze(0),
                         collision_fields=[CollisionField(cfg['obstacles'], tensor_args=dev)], tensor_args=dev, num_samples=S,
                         seed=seed, noise_particle_offset=p_off, noise_particles_global=P_glob, **sig)
    pl = make(99)
    means0 = pl._particle_means.clone()
    desc = pl._noise.desc()
    traj = pl.optimize(opt_iters=1)                      # no noise argument: drawn inside K1
    eps = pl._sample_dist.replay_noise(desc, S)
    ref = check.stoch_gpmp_subset(cfg, H, sig, means0, pl._sample_dist.scale_tril, pl.Sigma_inv, eps, range(P))
    amp = float((ref['samples'] - means0.cpu().unsqueeze(1)).abs().max())
    assert_close(pl.state_samples, ref['samples'], rtol=1e-5, atol=1e-5 * amp + 1e-7, what='samples')
    assert_close(pl.costs, ref['costs'], rtol=1e-5, what='costs')
    assert torch.equal(pl.free_flags.view(P, S).bool().cpu(), ref['free'])
    assert torch.equal(pl.costs.argmin(1).cpu(), ref['costs'].argmin(1))
    ref64 = check.stoch_gpmp_subset(cfg, H, sig, means0, pl._sample_dist.scale_tril, pl.Sigma_inv, eps, range(P), torch.float64)
    check.assert_not_worse_than_fp32(traj - means0, ref['means'] - means0.cpu(), ref64['means'] - means0.cpu().double(), 'mean update')
    # same seed => identical run (including the initial particles); the particles split over two "ranks" => identical too
    pl2 = make(99)
    assert torch.equal(pl2._particle_means, means0)
    assert torch.equal(pl2.optimize(opt_iters=1), traj)
    halves = []
    for r in range(2):
        ph = make(99, p_off=r * P // 2, P_loc=P // 2, P_glob=P)
        ph._particle_means.copy_(means0[r * P // 2:(r + 1) * P // 2])
        halves.append(ph.optimize(opt_iters=1))
    assert torch.equal(torch.cat(halves, dim=0), traj), 'result must not depend on how particles are sharded'
    # the next iteration draws fresh noise
    s0 = pl.state_samples.clone()
    pl._particle_means.copy_(means0)
    pl.optimize(opt_iters=1)
    assert not torch.equal(pl.state_samples, s0)
